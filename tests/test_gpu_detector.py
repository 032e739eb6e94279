"""The drop-in Detector (models/decoder.py:21-192) on the B200 captioner: one 'fact' and one 'senti' RL iteration with
stand-in sentiment models that have the reference modules' interfaces, plus Detector.sample."""
import pytest
import torch
import torch.nn as nn

from insenticap_model_b200 import synthetic as syn
from insenticap_model_b200.detector import Detector

pytestmark = pytest.mark.gpu
V, B, T = 300, 6, 8


class _StubSentiDetector(nn.Module):
    """sample(att_feats, threshold) -> (labels, senti_features, sentiment names, scores) like SentimentDetector.sample."""

    def __init__(self, cats):
        super().__init__()
        self.cats = cats

    def sample(self, features, senti_threshold=0):
        n = features.shape[0]
        labels = (torch.arange(n, device=features.device) % 3).long()
        return labels, None, [self.cats[int(i)] for i in labels], torch.ones(n, device=features.device)


class _StubSentCls(nn.Module):
    """forward(seqs, lengths) -> (pred [B,3], per-word weights [B, max(lengths)]) like SentenceSentimentClassifier."""

    def __init__(self):
        super().__init__()
        self.emb = nn.Embedding(V, 3)

    def forward(self, seqs, lengths):
        lens = torch.as_tensor([max(int(x), 1) for x in lengths], device=seqs.device)
        max_len = int(lens.max())
        mask = (torch.arange(max_len, device=seqs.device).unsqueeze(0) < lens.unsqueeze(1)).float()
        e = self.emb(seqs[:, :max_len]) * mask.unsqueeze(-1)
        return e.sum(1), mask / lens.unsqueeze(1)


def _detector():
    torch.manual_seed(0)
    d = Detector(syn.make_vocab(V), T, syn.SENTIMENT_CATEGORIES, {"cap_lr": 4e-4}, dict(syn.DEFAULT_SETTINGS),
                 senti_detector=_StubSentiDetector(syn.SENTIMENT_CATEGORIES), sent_senti_cls=_StubSentCls())
    d.captioner.load_state_dict(syn.synthetic_state_dict(V, 1))
    return d.cuda()


def _batches():
    fc, att, cpts, sentis, labels = syn.synthetic_inputs(B, V, seed=5)
    g = torch.Generator().manual_seed(6)
    caps = torch.randint(4, V, (B, T + 1), generator=g)
    caps[:, 0] = 1
    lengths = [T] * B
    refs = syn.synthetic_references(B, V, 5, seed=3)
    fns = ["img%d" % i for i in range(B)]
    gts = {fn: refs[i] for i, fn in enumerate(fns)}
    fact = [(fns, fc, att, (caps, lengths), cpts, sentis, gts)]
    senti = [(fns, fc, att, cpts, sentis, labels)]
    scs = [((caps, lengths), cpts, sentis, labels)]
    return fact, senti, scs, gts


def test_detector_fact_and_senti_iterations_and_sample():
    d = _detector()
    fact, senti, scs, gts = _batches()
    d.set_ciderd_scorer({"train": gts})
    before = {k: v.detach().clone() for k, v in d.captioner.named_parameters()}
    out = d((fact, scs), "fact", training=True)
    assert set(out) == {"da_loss", "fact_reward", "cls_reward", "all_rewards", "cap_loss", "xe_loss", "seq2seq_loss"}
    assert all(torch.isfinite(torch.tensor(v)) for v in out.values())
    moved = max(float((p.detach() - before[k]).abs().max()) for k, p in d.captioner.named_parameters())
    assert 0 < moved <= 4e-4 * 1.01  # one clamp + Adam step
    out2 = d((senti, scs), "senti", training=True)
    assert "fact_reward" not in out2 and "xe_loss" not in out2 and "seq2seq_loss" in out2
    snap = {k: v.detach().clone() for k, v in d.captioner.named_parameters()}
    out3 = d((fact, scs), "fact", training=False)  # evaluation: no optimizer step, no seq2seq pass
    assert "seq2seq_loss" not in out3
    assert all(torch.equal(p.detach(), snap[k]) for k, p in d.captioner.named_parameters())
    fc, att, cpts, sentis, labels = syn.synthetic_inputs(1, V, seed=9)
    caps, sentiments = d.sample(fc[0].cuda(), att[0].cuda(), sentis[0].cuda(), beam_size=3)
    assert len(caps) == 3 and isinstance(caps[0], str) and sentiments == ["positive"]


def test_detector_with_on_device_sentiment_models_and_cls_reward():
    """The whole Detector on libisc_b200.so: image sentiment detector, captioner and sentence sentiment classifier; the
    classifier reward (self_critical/utils.py:120-151) checked against the CPU oracle of the classifier."""
    import numpy as np
    from insenticap_model_b200.detector import get_cls_reward
    from insenticap_model_b200.sent_senti_cls import SentenceSentimentClassifier
    from insenticap_model_b200.sentiment_detector import SentimentDetector
    from oracle import sentcls_oracle as CO

    settings = dict(syn.DEFAULT_SETTINGS, sentiment_convs_num=2, sentiment_fcs_num=2)
    sd = SentimentDetector(syn.SENTIMENT_CATEGORIES, settings)
    sd.load_state_dict(syn.senti_detector_state_dict(0))
    sc = SentenceSentimentClassifier(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, settings)
    sc.load_state_dict(syn.sent_cls_state_dict(V, 0))
    torch.manual_seed(0)
    d = Detector(syn.make_vocab(V), T, syn.SENTIMENT_CATEGORIES, {"cap_lr": 4e-4}, settings, senti_detector=sd, sent_senti_cls=sc)
    d.captioner.load_state_dict(syn.synthetic_state_dict(V, 1))
    d = d.cuda()
    fact, senti, scs, gts = _batches()
    d.set_ciderd_scorer({"train": gts})
    out = d((fact, scs), "fact", training=True)
    assert all(torch.isfinite(torch.tensor(v)) for v in out.values()) and "cls_reward" in out
    out2 = d((senti, scs), "senti", training=True)
    assert all(torch.isfinite(torch.tensor(v)) for v in out2.values())

    seqs, lengths = syn.sent_cls_inputs(B, V, max_len=T, seed=77)
    masks = (torch.arange(T)[None, :] < torch.tensor(lengths)[:, None]).float()
    labels = torch.arange(B) % 3
    got = get_cls_reward(seqs.cuda(), masks.cuda(), None, None, labels.cuda(), sc)
    with torch.no_grad():
        pred, w = CO.forward(syn.sent_cls_state_dict(V, 0), seqs, lengths)
    want = torch.nn.functional.pad((pred.argmax(-1) == labels).float()[:, None] * w, (0, T - w.shape[1]))
    assert tuple(got.shape) == (B, T)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-5)


def _ref_config_detector(golden):
    """Our Detector in the configuration tests/golden/make_golden.py::detector_goldens ran the REFERENCE Detector in."""
    from insenticap_model_b200.sent_senti_cls import SentenceSentimentClassifier
    from insenticap_model_b200.sentiment_detector import SentimentDetector
    Vg, Tg = 64, 12
    cats = [str(c) for c in golden["categories"]]
    settings = dict(syn.DEFAULT_SETTINGS, sentiment_convs_num=2, sentiment_fcs_num=2)
    sd = SentimentDetector(cats, settings)
    sd.load_state_dict(syn.senti_detector_state_dict(0))
    sc = SentenceSentimentClassifier(syn.make_vocab(Vg), cats, settings)
    sc.load_state_dict(syn.sent_cls_state_dict(Vg, int(golden["cls_seed"])))
    torch.manual_seed(0)
    d = Detector(syn.make_vocab(Vg), Tg, cats, {"cap_lr": 4e-4}, settings, senti_detector=sd, sent_senti_cls=sc)
    d.captioner.load_state_dict(syn.synthetic_state_dict(Vg, 0, eos_heavy=True))
    return d.cuda(), Vg, Tg


def test_detector_matches_reference_detector_golden():
    """Detector.forward(training=False) and Detector.sample against the REFERENCE Detector (models/decoder.py:52-192) with
    the reference's own SentimentDetector / SentenceSentimentClassifier, on the same synthetic batches, same weights and
    the same injected Gumbel noise for the sampled pass: sampled and greedy ids exact, every reported loss / reward equal."""
    import os
    import numpy as np
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "detector_golden.npz"))
    d, Vg, Tg = _ref_config_detector(golden)
    Bg = 6
    fact, senti, scs, gts = syn.detector_batches(Bg, Vg, Tg)
    d.set_ciderd_scorer({"train": gts})
    d.sample_noise = syn.gumbel_noise(Tg, Bg, Vg, seed=21).cuda()
    for tag, data, dtype in (("fact", (fact, scs), "fact"), ("senti", (senti, scs), "senti")):
        with torch.no_grad():
            out = d(data, dtype, training=False)
        sample, greedy = d.last_captions
        assert np.array_equal(sample.cpu().numpy(), golden[tag + "_sample_seq"]), tag
        assert np.array_equal(greedy.cpu().numpy(), golden[tag + "_greedy_seq"]), tag
        keys = [k[len(tag) + 1:] for k in golden.files if k.startswith(tag + "_") and golden[k].shape == ()]
        assert set(keys) == set(out), (keys, sorted(out))
        for k in keys:
            np.testing.assert_allclose(out[k], float(golden[tag + "_" + k]), rtol=2e-3, atol=2e-5, err_msg=tag + " " + k)
    fc, att, cpts, sentis, labels = syn.synthetic_inputs(2, Vg, seed=9)
    for i in range(2):
        caps, sentiments = d.sample(fc[i].cuda(), att[i].cuda(), sentis[i].cuda(), beam_size=3)
        assert caps == [str(c) for c in golden["sample_captions"][i]]
        assert sentiments == [str(s) for s in golden["sample_sentiments"][i]]


def test_detector_iteration_has_one_host_sync():
    """The RL iteration keeps ids, rewards and losses on the device: the only torch-level host synchronisation of a
    Detector.forward call is the ONE stacked read of the loss sums after the loop (the reference does seven per iteration)."""
    import os
    import warnings
    import numpy as np
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "detector_golden.npz"))
    d, Vg, Tg = _ref_config_detector(golden)
    fact, senti, scs, gts = syn.detector_batches(6, Vg, Tg)
    d.set_ciderd_scorer({"train": gts})
    fact = [tuple(x.cuda() if torch.is_tensor(x) else x for x in fact[0][:3]) + ((fact[0][3][0].cuda(), fact[0][3][1]),)
            + tuple(x.cuda() if torch.is_tensor(x) else x for x in fact[0][4:])] * 3  # three iterations, inputs resident
    with torch.no_grad():
        d((fact, scs), "fact", training=False)  # warm-up: packs weights, sizes workspaces
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("warn")
    try:
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            with torch.no_grad():
                d((fact, scs), "fact", training=False)
    finally:
        torch.cuda.set_sync_debug_mode("default")
    syncs = [str(x.message) for x in w if "synchroniz" in str(x.message).lower()]
    assert len(syncs) <= 1, syncs
