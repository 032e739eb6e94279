"""Deterministic synthetic weights and inputs for the caption-decode hot path.

Everything here is generated on the CPU from seeded ``torch.Generator`` streams, so the
golden-vector generator (tests/golden/make_golden.py, which runs the real reference in the
authoring container), the CPU oracle, the GPU parity tests and bench.py all see bit-identical
tensors without shipping 88 MB of weights or 1.6 MB/image of features in the repo.

Shapes and names follow the reference's ``Captioner.state_dict()``
(/root/reference/models/captioner.py:121-161); the init
distributions follow torch's defaults for the layers the reference uses (Linear and
LSTMCell: U(-1/sqrt(fan), 1/sqrt(fan)); Embedding: N(0,1) with the PAD row zeroed).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

# network sizes from /root/reference/opts.py:80-95
DEFAULT_SETTINGS = {
    "word_emb_dim": 512,
    "fc_feat_dim": 2048,
    "att_feat_dim": 2048,
    "feat_emb_dim": 512,
    "dropout_p": 0.5,
    "rnn_hid_dim": 512,
    "att_hid_dim": 512,
}
SENTIMENT_CATEGORIES = ["positive", "negative", "neutral"]  # /root/reference/opts.py:25
NUM_REGIONS = 196  # 14 x 14
NUM_SENTI_WORDS = 10  # /root/reference/opts.py:62
NUM_CONCEPTS = 5  # /root/reference/opts.py:61


def make_vocab(vocab_size: int) -> list[str]:
    """idx2word with the reference's special-token order (/root/reference/preprocess.py:276)."""
    assert vocab_size >= 5
    return ["<PAD>", "<SOS>", "<EOS>", "<UNK>"] + ["w%d" % i for i in range(vocab_size - 4)]


def param_specs(vocab_size: int, settings: dict | None = None, n_senti: int = 3):
    """(name, shape, fan) for the 40 tensors of ``Captioner.state_dict()`` in reference order."""
    s = dict(DEFAULT_SETTINGS if settings is None else settings)
    E, F, H, A = s["word_emb_dim"], s["feat_emb_dim"], s["rnn_hid_dim"], s["att_hid_dim"]
    Dfc, Datt = s["fc_feat_dim"], s["att_feat_dim"]
    specs = [
        ("word_embed.0.weight", (vocab_size, E), None),
        ("senti_label_embed.0.weight", (n_senti, E), None),
    ]

    def lin(name, out_f, in_f):
        specs.append((name + ".weight", (out_f, in_f), in_f))
        specs.append((name + ".bias", (out_f,), in_f))

    def lstm(name, in_f, hid):
        specs.append((name + ".weight_ih", (4 * hid, in_f), hid))
        specs.append((name + ".weight_hh", (4 * hid, hid), hid))
        specs.append((name + ".bias_ih", (4 * hid,), hid))
        specs.append((name + ".bias_hh", (4 * hid,), hid))

    lin("fc_embed.0", F, Dfc)
    lin("cpt2fc.0", F, E)
    lin("att_embed.0", F, Datt)
    lstm("att_lstm", H + F + E, H)
    lin("att2att.0", A, F)
    lin("senti2att.0", A, E)
    lin("attention.cont_att.h2att", A, H)
    lin("attention.cont_att.att_alpha", 1, A)
    lin("attention.senti_att.h2word", A, H)
    lin("attention.senti_att.label2word", A, E)
    lin("attention.senti_att.word_alpha", 1, A)
    lin("attention.h2att", A, H)
    lin("attention.cont2att", A, F)
    lin("attention.senti2att", A, F)
    lin("attention.att_alpha", 1, A)
    lstm("lang_lstm", H + F, H)
    lin("classifier", vocab_size, H)
    return specs


def synthetic_state_dict(vocab_size: int, seed: int = 0, settings: dict | None = None,
                         eos_heavy: bool = False) -> "OrderedDict[str, torch.Tensor]":
    """Random-init weights, one independent seeded stream per tensor.

    ``eos_heavy`` rescales the random weights so that EOS appears at varied positions
    (greedy lengths 1..16, beams finishing at different steps, EOS at t=0) and the
    finish / carry / early-stop logic is exercised: classifier.weight *= 5,
    word_embed *= 3, both LSTMs' weight matrices *= 12, classifier.bias[EOS] += 2.
    """
    sd = OrderedDict()
    for i, (name, shape, fan) in enumerate(param_specs(vocab_size, settings)):
        g = torch.Generator().manual_seed(seed * 1000 + i)
        if fan is None:
            t = torch.randn(shape, generator=g, dtype=torch.float32)
            if name == "word_embed.0.weight":
                t[0].zero_()  # padding_idx row
        else:
            bound = 1.0 / math.sqrt(fan)
            t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound
        sd[name] = t
    if eos_heavy:
        sd["classifier.weight"] = sd["classifier.weight"] * 5.0
        sd["word_embed.0.weight"] = sd["word_embed.0.weight"] * 3.0
        for n in ("att_lstm", "lang_lstm"):
            sd[n + ".weight_ih"] = sd[n + ".weight_ih"] * 12.0
            sd[n + ".weight_hh"] = sd[n + ".weight_hh"] * 12.0
        sd["classifier.bias"][2] += 2.0
    return sd


def senti_detector_state_dict(seed: int = 0, settings: dict | None = None, n_cls: int = 3, n_convs: int = 2, n_fcs: int = 2,
                              spread: float = 150.0) -> "OrderedDict[str, torch.Tensor]":
    """Random-init weights of the image sentiment detector (/root/reference/models/sentiment_detector.py:6-28) under
    the reference's parameter names. ``spread`` scales the first output layer so that the class probabilities of
    different random images land on both sides of the 0.7 neutral threshold."""
    s = dict(DEFAULT_SETTINGS if settings is None else settings)
    sd = OrderedDict()
    c = s["fc_feat_dim"]
    i = 0

    def uni(shape, fan):
        nonlocal i
        g = torch.Generator().manual_seed(seed * 1000 + 500 + i)
        i += 1
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) / math.sqrt(fan)

    for k in range(n_convs):
        sd["convs.conv_%d.weight" % k] = uni((c // 2, c, 3, 3), c * 9)
        sd["convs.conv_%d.bias" % k] = uni((c // 2,), c * 9)
        c //= 2
    sd["senti_conv.weight"] = uni((n_cls, c, 1, 1), c)
    sd["senti_conv.bias"] = uni((n_cls,), c)
    for k in range(n_fcs):
        sd["output.%d.weight" % k] = uni((n_cls, n_cls), n_cls) * (spread if k == 0 else 1.0)
        sd["output.%d.bias" % k] = uni((n_cls,), n_cls)
    return sd


def sent_cls_state_dict(vocab_size: int, seed: int = 0, n_cls: int = 3, hid: int = 512) -> "OrderedDict[str, torch.Tensor]":
    """Random-init weights of the sentence sentiment classifier (/root/reference/models/sent_senti_cls.py:6-36) under the
    reference's parameter names (uniform +-1/sqrt(fan_in), embeddings N(0,1) with a zero <PAD> row, like torch's init)."""
    sd = OrderedDict()
    i = 0

    def uni(shape, fan):
        nonlocal i
        g = torch.Generator().manual_seed(seed * 1000 + 700 + i)
        i += 1
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) / math.sqrt(fan)

    emb = torch.randn(vocab_size, hid, generator=torch.Generator().manual_seed(seed * 1000 + 699))
    emb[0] = 0.0
    sd["word_embed.0.weight"] = emb
    sd["rnn.weight_ih_l0"] = uni((4 * hid, hid), hid)
    sd["rnn.weight_hh_l0"] = uni((4 * hid, hid), hid)
    sd["rnn.bias_ih_l0"] = uni((4 * hid,), hid)
    sd["rnn.bias_hh_l0"] = uni((4 * hid,), hid)
    for name, rows in (("excitation.0", hid), ("excitation.2", hid), ("sent_senti_cls.0", hid), ("sent_senti_cls.3", n_cls)):
        sd[name + ".weight"] = uni((rows, hid), hid)
        sd[name + ".bias"] = uni((rows,), hid)
    return sd


def sent_cls_inputs(batch: int, vocab_size: int, max_len: int = 16, seed: int = 31):
    """Captions int64 [B, max_len] (zero = <PAD> past each length) and ragged lengths [B] in 1..max_len, unsorted, with
    at least one caption of full length and one of length 1."""
    g = torch.Generator().manual_seed(seed)
    seqs = torch.randint(4, vocab_size, (batch, max_len), generator=g)
    lengths = torch.randint(1, max_len + 1, (batch,), generator=g)
    lengths[batch // 2] = max_len
    if batch > 1:
        lengths[0] = 1
    seqs = seqs * (torch.arange(max_len)[None, :] < lengths[:, None])
    return seqs, lengths.tolist()


def loader_items(n_images: int = 7, vocab_size: int = 50, feat_dim: int = 8, grid: int = 2, seed: int = 41):
    """Synthetic raw material of the reference's datasets (dataloader.py): per image a name, fc / att features, 1-4
    captions of 2-24 ids, 0-8 concept ids, 0-14 sentiment-word ids and a label; plus a sentiment corpus. Deterministic."""
    g = torch.Generator().manual_seed(seed)

    def ids(lo, hi):
        n = int(torch.randint(lo, hi + 1, (1,), generator=g))
        return torch.randint(4, vocab_size, (n,), generator=g).tolist()

    names = ["img_%03d.jpg" % i for i in range(n_images)]
    fc = torch.rand(n_images, feat_dim, generator=g)
    att = torch.rand(n_images, grid, grid, feat_dim, generator=g)
    captions = {fn: [[1] + ids(1, 22) + [2] for _ in range(int(torch.randint(1, 5, (1,), generator=g)))] for fn in names}
    concepts = {fn: ids(0, 8) for fn in names}
    sentiments = {fn: ids(0, 14) for fn in names}
    labels = [(fn, int(torch.randint(0, 3, (1,), generator=g))) for fn in names]
    corpus = [([1] + ids(1, 22) + [2], ids(0, 8), ids(0, 14), int(torch.randint(0, 3, (1,), generator=g))) for _ in range(9)]
    return {"names": names, "fc": fc, "att": att, "captions": captions, "concepts": concepts, "sentiments": sentiments,
            "labels": labels, "corpus": corpus}


def senti_detector_inputs(batch: int = 6, seed: int = 21) -> torch.Tensor:
    """Region features [B,14,14,2048] for the sentiment-detector tests: random features times a per-image signed scale, so
    that the winning probabilities land on both sides of the 0.7 threshold."""
    _, att, _, _, _ = synthetic_inputs(batch, 100, seed=seed)
    scale = torch.tensor([-3.0, -1.0, -0.3, 0.3, 1.0, 3.0])
    return att * scale[torch.arange(batch) % 6].view(batch, 1, 1, 1)


def synthetic_inputs(batch: int, vocab_size: int, seed: int = 1, settings: dict | None = None,
                     labels: str | int = "cycle"):
    """fc_feats, att_feats, cpt_words, senti_words, senti_labels.

    Draw order is fixed (fc, att, cpts, sentis) from ONE generator, as in the survey anchors.
    """
    s = dict(DEFAULT_SETTINGS if settings is None else settings)
    g = torch.Generator().manual_seed(seed)
    fc = torch.rand(batch, s["fc_feat_dim"], generator=g)
    att = torch.rand(batch, 14, 14, s["att_feat_dim"], generator=g)
    cpts = torch.randint(4, vocab_size, (batch, NUM_CONCEPTS), generator=g)
    sentis = torch.randint(4, vocab_size, (batch, NUM_SENTI_WORDS), generator=g)
    if labels == "cycle":
        lab = torch.arange(batch) % 3
    else:
        lab = torch.full((batch,), int(labels), dtype=torch.long)
    return fc, att, cpts, sentis, lab.long()


def synthetic_captions(batch: int, vocab_size: int, length: int = 17, seed: int = 2):
    """Teacher-forcing captions [B, length]: <SOS> then random words (cfg3 of BASELINE.json)."""
    g = torch.Generator().manual_seed(seed)
    caps = torch.randint(4, vocab_size, (batch, length), generator=g)
    caps[:, 0] = 1
    return caps


def synthetic_references(n_images: int, vocab_size: int, refs_per_image: int = 5, seed: int = 3):
    """Zipf-ish reference captions so n-grams repeat: list (per image) of list of id lists,
    each wrapped <SOS> ... <EOS> like the reference's ground truth (dataloader.py:64)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_images):
        caps = []
        for _ in range(refs_per_image):
            n = int(torch.randint(6, 15, (1,), generator=g))
            u = torch.rand(n, generator=g).clamp_(1e-6, 1.0)
            # inverse-CDF of a Pareto tail (alpha=1.3) -> heavy head of frequent ids
            ids = (u ** (-1.0 / 1.3)).floor().long() + 3
            ids.clamp_(max=vocab_size - 1)
            caps.append([1] + ids.tolist() + [2])
        out.append(caps)
    return out


def gumbel_noise(T: int, batch: int, vocab_size: int, seed: int = 21) -> torch.Tensor:
    """Standard Gumbel noise [T, B, V] for reproducible sampled decodes (argmax(log p + g) draws from p)."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(T, batch, vocab_size, generator=g).clamp_(1e-10, 1.0 - 1e-7)
    return -torch.log(-torch.log(u))


def detector_batches(batch: int, vocab_size: int, max_seq_len: int):
    """In-memory batches with the tuple layouts Detector.forward consumes (/root/reference/models/decoder.py:64-76,
    :149-158; the loaders' collate outputs, dataloader.py:70-109): one 'fact' batch, one 'senti' batch and one
    sentiment-corpus (seq2seq) batch, plus the ground-truth dict of the fact images."""
    fc, _, cpts, sentis, labels = synthetic_inputs(batch, vocab_size, seed=5)
    att = senti_detector_inputs(batch)  # signed per-image scales: the image sentiment detector then yields varied labels
    g = torch.Generator().manual_seed(6)
    caps = torch.randint(4, vocab_size, (batch, max_seq_len + 1), generator=g)
    caps[:, 0] = 1
    lengths = [max_seq_len] * batch
    refs = synthetic_references(batch, vocab_size, 5, seed=3)
    fns = ["img%d" % i for i in range(batch)]
    gts = {fn: refs[i] for i, fn in enumerate(fns)}
    fact = [(fns, fc, att, (caps, lengths), cpts, sentis, gts)]
    senti = [(fns, fc, att, cpts, sentis, labels)]
    scs = [((caps, lengths), cpts, sentis, labels)]
    return fact, senti, scs, gts
