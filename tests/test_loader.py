"""Input pipeline, CPU: the collate functions against outputs of the reference's own
(tests/golden/loader_golden.json, made by tests/golden/make_golden.py), fed both reference-style items (numpy features)
and lazy shard references; the native shard reader against an independent numpy parse of the file format."""
import json
import os
import random
import struct

import numpy as np
import pytest
import torch

from insenticap_model_b200 import dataloader as dl
from insenticap_model_b200 import synthetic as syn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loader_golden.json")
NAMES = ["caption", "rl_fact", "rl_senti", "senti_image", "concept", "senti_corpus_with_sentis", "senti_sents"]


def jsonable(x):
    if torch.is_tensor(x):
        return {"dtype": str(x.dtype), "shape": list(x.shape), "data": x.reshape(-1).tolist()}
    if isinstance(x, (tuple, list)):
        return [jsonable(y) for y in x]
    if isinstance(x, dict):
        return {k: jsonable(v) for k, v in x.items()}
    if isinstance(x, np.generic):
        return x.item()
    return x


@pytest.fixture(scope="module")
def shard(tmp_path_factory):
    it = syn.loader_items()
    path = str(tmp_path_factory.mktemp("shards") / "feats.iscf")
    dl.FeatureShard.write(path, it["names"][::-1], it["fc"].flip(0), it["att"].flip(0))  # record order != dataset order
    return dl.FeatureShard(path)


def _items(it, name, shard=None):
    fc, att = it["fc"].numpy(), it["att"].numpy()
    row = {fn: i for i, fn in enumerate(it["names"])}
    if shard is None:
        f = lambda fn: fc[row[fn]]
        a = lambda fn: att[row[fn]]
    else:
        f = lambda fn: dl._Lazy(shard, shard.index(fn), "fc")
        a = lambda fn: dl._Lazy(shard, shard.index(fn), "att")
    return {
        "caption": lambda: [(fn, f(fn), a(fn), caps, it["concepts"][fn]) for fn, caps in it["captions"].items()],
        "rl_fact": lambda: [(fn, caps, f(fn), a(fn), it["concepts"][fn], it["sentiments"][fn]) for fn, caps in it["captions"].items()],
        "rl_senti": lambda: [(fn, f(fn), a(fn), it["concepts"][fn], it["sentiments"][fn], lab) for fn, lab in it["labels"]],
        "senti_image": lambda: [(fn, a(fn), lab) for fn, lab in it["labels"]],
        "concept": lambda: [(fn, f(fn), np.eye(1, 20, k=row[fn], dtype=np.int16)[0]) for fn in it["names"]],
        "senti_corpus_with_sentis": lambda: list(it["corpus"]),
        "senti_sents": lambda: [(lab, np.array(cap)) for cap, _, _, lab in it["corpus"]],
    }[name]()


@pytest.mark.parametrize("lazy", [False, True])
@pytest.mark.parametrize("name", NAMES)
def test_collate_matches_reference_golden(name, lazy, shard):
    gold = json.load(open(GOLD))[name]
    it = syn.loader_items()
    random.seed(5)
    fn = dl.create_collate_fn(name, pad_index=0, max_seq_len=17, num_concepts=5, num_sentiments=10)
    got = json.loads(json.dumps(jsonable(fn(_items(it, name, shard if lazy else None)))))
    assert got == gold  # ids, lengths, names, ordering and the fp32 features, all exact


def test_unknown_collate_name_is_none_like_the_reference():
    assert dl.create_collate_fn("nope") is None


def _numpy_parse(path):
    """Independent reader of the documented layout (include/isc.h)."""
    raw = open(path, "rb").read()
    magic, version, dtype, d, l, n, names_bytes, data_off, rec = struct.unpack("<8sIIIIQQQQ", raw[:56])
    assert magic == b"ISCFEAT1" and version == 1
    names = raw[64:64 + names_bytes].split(b"\0")[:-1]
    assert len(names) == n
    np_dt = np.float32 if dtype == 0 else np.uint16
    recs = np.frombuffer(raw, dtype=np.uint8, offset=data_off).reshape(n, rec)
    elems = recs[:, :d * (1 + l) * np.dtype(np_dt).itemsize].copy().view(np_dt).reshape(n, (1 + l) * d)
    return [s.decode() for s in names], elems[:, :d], elems[:, d:].reshape(n, l, d), data_off, rec


@pytest.mark.parametrize("dtype", ["fp32", "bf16", "fp16"])
def test_shard_round_trip_against_numpy_parse(tmp_path, dtype):
    n, d, l = 37, 24, 9
    g = torch.Generator().manual_seed(3)
    fc = torch.randn(n, d, generator=g)
    att = torch.randn(n, 3, 3, d, generator=g)
    names = ["COCO_%06d.jpg" % (7 * i) for i in range(n)]
    path = dl.FeatureShard.write(str(tmp_path / "s.iscf"), names, fc, att, dtype=dtype)
    pn, pfc, patt, off, rec = _numpy_parse(path)
    assert pn == names and off % 256 == 0 and rec % 256 == 0
    sh = dl.FeatureShard(path)
    assert (len(sh), sh.feat_dim, sh.n_regions, sh.att_shape) == (n, d, l, (3, 3, d))
    idx = [5, 0, 36, 5, 5, 12]  # unordered, repeated
    for threads in (1, 4):
        gfc, gatt = sh.gather(idx, threads=threads, pin=False)
        if dtype == "fp32":
            assert torch.equal(gfc, fc[idx]) and torch.equal(gatt, att[idx])
            assert np.array_equal(gfc.numpy(), pfc[idx])
        else:
            tdt = torch.bfloat16 if dtype == "bf16" else torch.float16
            assert gfc.dtype == tdt
            assert torch.equal(gfc, fc[idx].to(tdt)) and torch.equal(gatt, att[idx].to(tdt))  # round-to-nearest-even
            assert np.array_equal(gfc.view(torch.int16).numpy().view(np.uint16), pfc[idx])
            assert np.array_equal(gatt.view(torch.int16).numpy().view(np.uint16).reshape(len(idx), l, d), patt[idx])
    only_att = sh.gather([1, 2], want_fc=False, pin=False)
    assert only_att[0] is None and tuple(only_att[1].shape) == (2, 3, 3, d)
    assert sh.index(names[12]) == 12 and sh.name(12) == names[12] and names[3] in sh and "nope" not in sh
    one_fc, one_att = sh[names[4]]
    assert one_fc.dtype == np.float32 and one_att.shape == (3, 3, d)
    with pytest.raises(KeyError):
        sh.index("missing.jpg")
    with pytest.raises(IndexError):
        sh.name(n)
    with pytest.raises(RuntimeError, match="out of range"):
        sh.gather([0, n], pin=False)
    assert tuple(sh.gather([], pin=False)[0].shape) == (0, d)  # empty batch


def test_shard_open_rejects_bad_files(tmp_path):
    good = dl.FeatureShard.write(str(tmp_path / "g.iscf"), ["a", "b"], torch.zeros(2, 8), torch.zeros(2, 4, 8))
    raw = open(good, "rb").read()
    for name, blob in (("magic", b"NOTASHRD" + raw[8:]), ("trunc", raw[:-300]), ("short", raw[:20])):
        p = str(tmp_path / (name + ".iscf"))
        open(p, "wb").write(blob)
        with pytest.raises(RuntimeError, match="shard"):
            dl.FeatureShard(p)
    with pytest.raises(RuntimeError, match="cannot open"):
        dl.FeatureShard(str(tmp_path / "absent.iscf"))
    with pytest.raises(ValueError):
        dl.FeatureShard.write(str(tmp_path / "x.iscf"), ["a", "a"], torch.zeros(2, 8), torch.zeros(2, 4, 8))


def test_loader_factories_cover_every_item(shard):
    """The get_*_dataloader factories (dataloader.py:277-370): same signatures, every dataset item exactly once."""
    it = syn.loader_items()
    p = shard.path
    n_caps = sum(len(c) for c in it["captions"].values())
    seen = 0
    for fns, fc, att, (caps, lengths), cpts in dl.get_caption_dataloader(p, p, it["captions"], it["concepts"], 0, 16, 5, 3, shuffle=False):
        assert fc.shape[0] == att.shape[0] == caps.shape[0] == len(fns) == len(lengths) and caps.shape[1] <= 17
        assert lengths == sorted(lengths, reverse=True) and tuple(att.shape[1:]) == (2, 2, 8)
        for i, fn in enumerate(fns):
            assert torch.equal(fc[i], it["fc"][it["names"].index(fn)])
        seen += len(fns)
    assert seen == n_caps
    random.seed(0)
    batches = list(dl.get_rl_fact_dataloader(shard, shard, it["captions"], it["concepts"], it["sentiments"], 0, 16, 5, 10, 4, shuffle=False))
    assert sum(len(b[0]) for b in batches) == len(it["names"]) and set(batches[0][6]) == set(batches[0][0])
    batches = list(dl.get_rl_senti_dataloader(p, p, it["concepts"], it["sentiments"], it["labels"], 0, 5, 10, 4, shuffle=False))
    assert [lab for b in batches for lab in b[5].tolist()] == [lab for _, lab in it["labels"]]
    assert tuple(batches[0][4].shape) == (4, 10) and tuple(batches[0][3].shape) == (4, 5)
    batches = list(dl.get_senti_image_dataloader(p, it["labels"], 5, shuffle=False))
    assert tuple(batches[0][1].shape) == (5, 2, 2, 8)
    batches = list(dl.get_concept_dataloader(p, {fn: [1, 3] for fn in it["names"]}, 6, 7, shuffle=False))
    assert batches[0][2].tolist() == [[0, 1, 0, 1, 0, 0]] * 7
    (caps, lengths), cpts, sentis, ids = next(iter(dl.get_senti_corpus_with_sentis_dataloader(it["corpus"], 0, 16, 5, 10, 9, shuffle=False)))
    assert caps.shape[0] == 9 and max(lengths) <= 16
    sents = [(lab, cap) for cap, _, _, lab in it["corpus"]]
    ids, (caps, lengths) = next(iter(dl.get_senti_sents_dataloader(sents, 0, 16, batch_size=9, num_workers=0, shuffle=False)))
    assert caps.shape == (9, min(16, max(len(c) for _, c in sents))) and lengths[0] == caps.shape[1]
    eager = dl.CaptionDataset(p, p, it["captions"], it["concepts"], lazy=False)[0]  # reference-style item: numpy features
    assert isinstance(eager[1], np.ndarray) and eager[2].shape == (2, 2, 8)


def test_page_locked_shard_collate_defers_its_copies_for_the_prefetcher(shard, monkeypatch):
    """A page-locked shard's collate issues its device copies itself — except inside DevicePrefetcher's collate thread
    (thread-local switch), where it leaves `_Deferred(shard, record indices, which)` placeholders for the copy thread:
    the record order, the rest of the batch and the placeholder's target are what the direct path would have copied."""
    it = syn.loader_items()
    fn = dl.create_collate_fn("rl_senti", pad_index=0, num_concepts=5, num_sentiments=10)
    items = _items(it, "rl_senti", shard)
    want = fn(_items(it, "rl_senti", shard))  # staged path: pinned / pageable host tensors
    calls = []
    monkeypatch.setattr(shard, "direct_device", torch.device("cpu"), raising=False)  # stands in for pin(): no GPU here
    monkeypatch.setattr(shard, "copy_to_device",
                        lambda idx, want_fc=True, want_att=True: calls.append((list(idx), want_fc, want_att)) or
                        shard.gather(idx, want_fc=want_fc, want_att=want_att, pin=False), raising=False)
    dl._tls.defer_copies = True
    try:
        got = fn(items)
    finally:
        dl._tls.defer_copies = False
    assert isinstance(got[1], dl._Deferred) and isinstance(got[2], dl._Deferred) and not calls  # nothing copied yet
    assert (got[1].which, got[2].which) == ("fc", "att") and got[1].idx == got[2].idx == [shard.index(f) for f in got[0]]
    assert got[0] == want[0] and all(torch.equal(a, b) for a, b in zip(got[3:], want[3:]))
    fc, att = got[1].resolve(), got[2].resolve()
    assert calls == [(got[1].idx, True, False), (got[1].idx, False, True)]
    assert torch.equal(fc, want[1]) and torch.equal(att, want[2])
    direct = fn(items)  # outside the prefetcher: copied at once
    assert torch.is_tensor(direct[1]) and torch.equal(direct[2], want[2]) and len(calls) == 4
