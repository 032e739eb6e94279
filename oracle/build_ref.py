"""Recipe for oracle/_ref — TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

The reference (ezeli/InSentiCap_model) is a pure-Python package: there is nothing to compile with gcc. What
CAN be built from its sources "where they lie" is CPython bytecode: this script byte-compiles the reference
modules on the hot path (models/captioner.py, models/decoder.py and the modules they import, the vendored
CIDEr-D scorer and self_critical/utils.py) from /root/reference into sourceless bytecode files under
``oracle/_ref/`` (git-ignored, NOT gpurun-ignored: like our own built ``.so`` it travels to the GPU box, where
/root/reference does not exist). The files carry the standard ``.pyc`` format under the extension ``.refbc`` (the
snapshot sync to the GPU box skips ``*.pyc``); ``import_reference()`` registers an import hook for that extension on
the ``oracle/_ref`` directory only. No reference SOURCE file is copied into the repo.

    python -m oracle.build_ref            # writes oracle/_ref/**.pyc, prints the module list

Users: ``bench.py --impl reference`` / ``cpu_baseline`` (kind "reference": the UNMODIFIED reference's own
``Captioner.sample`` timed on the host cores) and tests that cross-check the oracle port against it when the
directory exists. The product package never imports anything from here.
"""
from __future__ import annotations

import os
import py_compile
import sys

EXT = ".refbc"
HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("ISC_REFERENCE_SRC", "/root/reference")
OUT = os.path.join(HERE, "_ref")

# modules on the path and the package __init__ files needed to import them
MODULES = [
    "models/__init__.py",
    "models/captioner.py",
    "models/decoder.py",
    "models/sentiment_detector.py",
    "models/sent_senti_cls.py",
    "self_critical/__init__.py",
    "self_critical/utils.py",
    "self_critical/cider/__init__.py",
    "self_critical/cider/pyciderevalcap/__init__.py",
    "self_critical/cider/pyciderevalcap/ciderD/__init__.py",
    "self_critical/cider/pyciderevalcap/ciderD/ciderD.py",
    "self_critical/cider/pyciderevalcap/ciderD/ciderD_scorer.py",
    "self_critical/bleu/__init__.py",
    "self_critical/bleu/bleu.py",
    "self_critical/bleu/bleu_scorer.py",
]


def build(verbose: bool = False) -> str | None:
    """Byte-compile the reference modules into oracle/_ref. Returns the directory, or None when the reference
    sources are not present (the GPU box: the prebuilt files shipped with the snapshot are used as they are)."""
    if not os.path.isdir(REF_SRC):
        return OUT if os.path.isdir(OUT) else None
    done = []
    for rel in MODULES:
        src = os.path.join(REF_SRC, rel)
        if not os.path.exists(src):
            if rel.endswith("__init__.py"):  # namespace-style directory in the reference: make it a package
                dst = os.path.join(OUT, rel[:-3] + EXT)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                empty = os.path.join(OUT, ".empty.py")
                open(empty, "w").close()
                py_compile.compile(empty, cfile=dst, doraise=True)
                os.remove(empty)
                done.append(rel + " (empty)")
            continue
        dst = os.path.join(OUT, rel[:-3] + EXT)  # sourceless import: pkg/mod.refbc where mod.py would be
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile=rel, doraise=True)
        done.append(rel)
    with open(os.path.join(OUT, "MANIFEST.txt"), "w") as f:
        f.write("# bytecode built by oracle/build_ref.py from %s with CPython %s\n" % (REF_SRC, sys.version.split()[0]))
        f.write("\n".join(done) + "\n")
    if verbose:
        print("\n".join(done))
    return OUT


def available() -> bool:
    return os.path.exists(os.path.join(OUT, "models", "captioner" + EXT))


def _install_hook():
    """Import hook: directories under oracle/_ref resolve modules from ``*.refbc`` (pyc-format) files."""
    from importlib.machinery import FileFinder, SourcelessFileLoader
    base = FileFinder.path_hook((SourcelessFileLoader, [EXT]))

    def hook(path):
        if not os.path.abspath(path).startswith(OUT):
            raise ImportError("not an oracle/_ref directory")
        return base(path)

    if not any(getattr(h, "_isc_ref_hook", False) for h in sys.path_hooks):
        hook._isc_ref_hook = True
        sys.path_hooks.insert(0, hook)
        sys.path_importer_cache.pop(OUT, None)


def import_reference():
    """Put oracle/_ref on sys.path and return the reference's ``models.captioner`` module (unmodified bytecode)."""
    if not available():
        raise ImportError("oracle/_ref is not built (python -m oracle.build_ref, needs /root/reference)")
    _install_hook()
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    import importlib
    importlib.invalidate_caches()
    return importlib.import_module("models.captioner")


if __name__ == "__main__":
    out = build(verbose=True)
    print(out if out else "reference sources not found at %s" % REF_SRC)
