"""Input pipeline on the GPU: shard -> pinned gather -> prefetched H2D -> beam search gives
the captions of the directly supplied tensors; a bf16 shard in precision="bf16" is bit-identical to fp32 inputs there."""
import pytest
import torch

from insenticap_model_b200 import dataloader as dl
from insenticap_model_b200 import synthetic as syn
from insenticap_model_b200.captioner import Captioner

pytestmark = pytest.mark.gpu
V, N = 400, 20


def _captioner(precision):
    m = Captioner(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS), precision=precision)
    m.load_state_dict(syn.synthetic_state_dict(V, 2))
    return m.cuda().eval()


@pytest.fixture(scope="module")
def corpus(tmp_path_factory):
    fc, att, cpts, sentis, labels = syn.synthetic_inputs(N, V, seed=12)
    names = ["img%03d" % i for i in range(N)]
    d = tmp_path_factory.mktemp("shards")
    paths = {k: dl.FeatureShard.write(str(d / (k + ".iscf")), names, fc, att, dtype=k) for k in ("fp32", "bf16", "fp16")}
    meta = dict(names=names, fc=fc, att=att, concepts={fn: cpts[i].tolist() for i, fn in enumerate(names)},
                sentiments={fn: sentis[i].tolist() for i, fn in enumerate(names)},
                labels=[(fn, int(labels[i])) for i, fn in enumerate(names)], sentis=sentis, lab=labels)
    return paths, meta


def test_prefetched_shard_batches_decode_like_direct_tensors(corpus):
    paths, meta = corpus
    m = _captioner("bf16x3")
    with torch.no_grad():
        want = m.beam_search(meta["fc"].cuda(), meta["att"].cuda(), meta["sentis"].cuda(), meta["lab"].cuda(), 3, 1, 16)
        want = [x.clone() for x in want]
        loader = dl.get_rl_senti_dataloader(paths["fp32"], paths["fp32"], meta["concepts"], meta["sentiments"], meta["labels"],
                                            0, 5, 10, batch_size=8, shuffle=False)
        seen = 0
        for fns, fc, att, cpts, sentis, labels in dl.DevicePrefetcher(loader, "cuda:0", depth=2):
            assert fc.is_cuda and att.is_cuda and att.dtype == torch.float32 and tuple(att.shape[1:]) == (14, 14, 2048)
            assert list(fns) == meta["names"][seen:seen + len(fns)]
            tok, sc, ln = m.beam_search(fc, att, sentis, labels, 3, 1, 16)
            assert torch.equal(tok, want[0][seen:seen + len(fns)]) and torch.equal(ln, want[2][seen:seen + len(fns)])
            assert torch.equal(sc, want[1][seen:seen + len(fns)])
            seen += len(fns)
        assert seen == N


def test_bf16_shard_is_bit_identical_in_bf16_mode_and_rejected_elsewhere(corpus):
    paths, meta = corpus
    m = _captioner("bf16")
    sh = dl.FeatureShard(paths["bf16"])
    fc16, att16 = sh.gather(range(N))
    assert fc16.dtype == torch.bfloat16 and fc16.is_pinned()
    with torch.no_grad():
        want = [x.clone() for x in m.beam_search(meta["fc"].cuda(), meta["att"].cuda(), meta["sentis"].cuda(), meta["lab"].cuda(), 3, 1, 16)]
        got = m.beam_search(fc16.cuda(), att16.cuda(), meta["sentis"].cuda(), meta["lab"].cuda(), 3, 1, 16)
        assert all(torch.equal(a, b) for a, b in zip(got, want))
        # host (pinned) bf16 tensors through the pipelined host path: same captions, half the H2D bytes
        got_h = m.beam_search(fc16, att16, meta["sentis"], meta["lab"], 3, 1, 16, host_chunk=8)
        assert all(torch.equal(a.cuda(), b) for a, b in zip(got_h, want))
        # greedy decode (forward_rl) through the same prologue
        seq, lp, mask = m(fc16.cuda(), att16.cuda(), None, meta["sentis"].cuda(), meta["lab"].cuda(), 16, sample_max=1, mode="rl")
        seq2, lp2, _ = m(meta["fc"].cuda(), meta["att"].cuda(), None, meta["sentis"].cuda(), meta["lab"].cuda(), 16, sample_max=1, mode="rl")
        assert torch.equal(seq, seq2) and torch.equal(lp, lp2)
        # the token-exact mode computes on fp32 inputs: bf16 tensors are widened, not reinterpreted
        x3 = _captioner("bf16x3")
        a = x3.beam_search(fc16.cuda(), att16.cuda(), meta["sentis"].cuda(), meta["lab"].cuda(), 3, 1, 16)
        b = x3.beam_search(fc16.float().cuda(), att16.float().cuda(), meta["sentis"].cuda(), meta["lab"].cuda(), 3, 1, 16)
        assert all(torch.equal(p, q) for p, q in zip(a, b))


def test_pinned_shard_collates_straight_into_device_tensors(corpus):
    """isc_shard_pin + isc_shard_copy_to_device: records DMA'd from the page cache, no staging copy; same bytes."""
    paths, meta = corpus
    for kind in ("fp32", "bf16"):
        sh = dl.FeatureShard(paths[kind])
        assert sh.pin("cuda:0") and sh.name(3) == meta["names"][3] and sh.index(meta["names"][5]) == 5
        loader = dl.get_rl_senti_dataloader(sh, sh, meta["concepts"], meta["sentiments"], meta["labels"], 0, 5, 10,
                                            batch_size=7, shuffle=False)
        seen = 0
        for fns, fc, att, cpts, sentis, labels in dl.DevicePrefetcher(loader, "cuda:0", depth=2):
            n = len(fns)
            want_fc, want_att = meta["fc"][seen:seen + n], meta["att"][seen:seen + n]
            if kind == "bf16":
                want_fc, want_att = want_fc.bfloat16(), want_att.bfloat16()
            assert fc.is_cuda and torch.equal(fc.cpu(), want_fc) and torch.equal(att.cpu(), want_att)
            assert labels.is_cuda and labels.tolist() == [l for _, l in meta["labels"][seen:seen + n]]
            seen += n
        assert seen == N


def test_fp16_shard_is_the_fp32_computation_on_the_stored_values(corpus):
    """A fp16 feature shard halves the host->device bytes in EVERY precision: fp16 values are fp32 values (isc_expand_f16 is
    exact) and fit the split-bf16 GEMM operands exactly, so decoding the shard's tensors — device tensors, or pinned host
    tensors through the pipelined host path — is bit-identical to decoding their fp32 widening, and token-exact against the
    CPU oracle run on those same values."""
    import numpy as np
    from oracle import captioner_oracle as O
    paths, meta = corpus
    sh = dl.FeatureShard(paths["fp16"])
    fc16, att16 = sh.gather(range(N))
    assert fc16.dtype == torch.float16 and fc16.is_pinned() and torch.equal(fc16, meta["fc"].half())
    sentis, lab = meta["sentis"], meta["lab"]
    for precision in ("bf16x3", "bf16"):
        m = _captioner(precision)
        with torch.no_grad():
            want = [x.clone() for x in m.beam_search(fc16.float().cuda(), att16.float().cuda(), sentis.cuda(), lab.cuda(), 3, 1, 16)]
            got = m.beam_search(fc16.cuda(), att16.cuda(), sentis.cuda(), lab.cuda(), 3, 1, 16)
            assert all(torch.equal(a, b) for a, b in zip(got, want))
            got_h = m.beam_search(fc16, att16, sentis, lab, 3, 1, 16, host_chunk=8)
            assert all(torch.equal(a.cuda(), b) for a, b in zip(got_h, want))
    m = _captioner("bf16x3")
    p = syn.synthetic_state_dict(V, 2)
    with torch.no_grad():
        got = m.beam_search(fc16.cuda(), att16.cuda(), sentis.cuda(), lab.cuda(), 3, 1, 16)
        f = O.prologue(p, fc16.float(), att16.float(), None, sentis, lab)
        tk_o, sc_o, ln_o, margin = O.beam_search(p, f, N, 3, 1, 16, return_margins=True)
    if margin > 1e-5:
        assert np.array_equal(got[0].cpu().numpy(), tk_o.numpy())
        np.testing.assert_allclose(got[1].cpu().numpy(), sc_o.numpy(), atol=2e-4)


def test_prefetcher_stops_early_and_surfaces_loader_errors(corpus):
    """The two worker threads (collate, copy) end when the consumer leaves the loop early, and an exception raised inside
    the loader (here: by the collate) reaches the consumer instead of hanging it."""
    import threading
    paths, meta = corpus
    sh = dl.FeatureShard(paths["fp16"])
    loader = dl.get_rl_senti_dataloader(sh, sh, meta["concepts"], meta["sentiments"], meta["labels"], 0, 5, 10, batch_size=4,
                                        shuffle=False)
    before = threading.active_count()
    for i, batch in enumerate(dl.DevicePrefetcher(loader, "cuda:0", depth=2)):
        assert batch[2].is_cuda and batch[2].dtype == torch.float16
        if i == 1:
            break
    assert threading.active_count() <= before  # both workers joined

    class Broken(torch.utils.data.Dataset):
        def __len__(self):
            return 8

        def __getitem__(self, i):
            if i == 5:
                raise ValueError("broken item")
            return torch.zeros(3)

    bad = torch.utils.data.DataLoader(Broken(), batch_size=2)
    with pytest.raises(ValueError, match="broken item"):
        for _ in dl.DevicePrefetcher(bad, "cuda:0", depth=2):
            pass
