"""The C-ABI library builds, loads, and exports exactly what include/isc.h declares. CPU only:
no compute entry point is called (they need a B200)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from insenticap_model_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "isc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(isc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    from insenticap_model_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libisc_b200.so does not export %s" % name
    assert sorted(_lib.SIGNATURES) == declared, "ctypes signature table out of sync with include/isc.h"


def test_version_and_size_queries(lib):
    from insenticap_model_b200 import _lib
    assert b"sm_100a" in lib.isc_version()
    d = _lib.Dims(10000, 512, 2048, 196, 11, 3, 0, 1, 2, 3)
    fp32 = lib.isc_packed_weights_bytes(ctypes.byref(d), _lib.PREC_FP32)
    x3 = lib.isc_packed_weights_bytes(ctypes.byref(d), _lib.PREC_BF16X3)
    # 22.06 M parameters: fused fp32 copy ~ 88 MB, plus two bf16 planes of the matrices
    assert 80e6 < fp32 < 120e6 and x3 > 1.8 * fp32 * 0.9
    assert lib.isc_decode_workspace_bytes(ctypes.byref(d), _lib.PREC_BF16X3, 3072) > 3072 * 10000 * 4
    assert lib.isc_prologue_workspace_bytes(ctypes.byref(d), _lib.PREC_BF16, 8) > 0
    bad = _lib.Dims(10000, 256, 2048, 196, 11, 3, 0, 1, 2, 3)  # hidden != 512 is not compiled in
    assert lib.isc_packed_weights_bytes(ctypes.byref(bad), _lib.PREC_FP32) == 0
    assert b"hidden" in lib.isc_last_error()
    assert lib.isc_cider_table_bytes(1024) == 1024 * 12 and lib.isc_cider_table_bytes(1000) == 0


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """The GEMM really is tcgen05 + TMA: UTCHMMA / UTMALDG / LDTM in the sm_100a SASS."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    from insenticap_model_b200 import build
    out = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in out, mnemonic
    assert "sm_100a" in out
