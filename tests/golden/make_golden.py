"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the authoring container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/decode_golden.npz and tests/golden/cider_golden.npz. Inputs and weights
are NOT stored: they are regenerated bit-identically from insenticap_model_b200.synthetic
(seeded CPU generators); a checksum of each is stored to catch RNG drift.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from insenticap_model_b200 import synthetic as syn  # noqa: E402
from models.captioner import Captioner as RefCaptioner  # noqa: E402
from self_critical.utils import get_ciderd_scorer, get_self_critical_reward  # noqa: E402
from self_critical.cider.pyciderevalcap.ciderD.ciderD_scorer import precook  # noqa: E402

T = 16


def build_ref(V, seed, eos_heavy=False):
    m = RefCaptioner(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS))
    m.load_state_dict(syn.synthetic_state_dict(V, seed, eos_heavy=eos_heavy))
    return m.eval()


def words_to_ids(caption, T):
    ids = [int(w[1:]) + 4 for w in caption.split()] if caption else []
    if len(ids) < T:
        ids.append(2)
    return ids


def ref_beam(m, fc, att, sentis, labels, K, constraint, xe_mode=False):
    B = fc.shape[0]
    toks = np.zeros((B, K, T), dtype=np.int64)
    lens = np.zeros((B, K), dtype=np.int32)
    scores = np.zeros((B, K), dtype=np.float64)
    with torch.no_grad():
        for i in range(B):
            if xe_mode:
                caps, sc = m.sample(fc[i], att[i], beam_size=K, decoding_constraint=constraint, max_seq_len=T)
            else:
                caps, sc = m.sample(fc[i], att[i], sentis[i], labels[i:i + 1], beam_size=K,
                                    decoding_constraint=constraint, max_seq_len=T)
            for k in range(K):
                ids = words_to_ids(caps[k], T)
                toks[i, k, :len(ids)] = ids
                lens[i, k] = len(ids)
                scores[i, k] = sc[k]
    return toks, lens, scores


def checksum(t):
    return float(t.double().sum())


def decode_goldens():
    out = {}
    # ---- cfg1: V=10000, B=8 -------------------------------------------------------------
    V, B = 10000, 8
    m = build_ref(V, 0)
    fc, att, cpts, sentis, labels = syn.synthetic_inputs(B, V, seed=1)
    out["cfg1_checksum_inputs"] = np.array([checksum(fc), checksum(att), checksum(cpts), checksum(sentis)])
    out["cfg1_checksum_weights"] = np.array([checksum(v) for v in m.state_dict().values()])
    with torch.no_grad():
        seq, lp, mask = m(fc, att, cpts, sentis, labels, T, 1, mode="rl")
        out["cfg1_greedy_seq"], out["cfg1_greedy_lp"], out["cfg1_greedy_mask"] = seq.numpy(), lp.numpy(), mask.numpy()
        out["cfg1_fc_embedded"] = m.fc_feats.numpy()
        out["cfg1_cpt_feats"] = m.cpt_feats.numpy()
        out["cfg1_cont_weights_sum"] = m.cont_weights.double().sum(0).numpy()  # [T*196]
        out["cfg1_senti_weights"] = m.senti_weights.numpy()  # [B, T*11]
        out["cfg1_gate_weights"] = m.cont_senti_weights.numpy()  # [B, T]
    tk, ln, sc = ref_beam(m, fc, att, sentis, labels, 3, 1)
    out["cfg1_beam3_tokens"], out["cfg1_beam3_lens"], out["cfg1_beam3_scores"] = tk, ln, sc
    tk, ln, sc = ref_beam(m, fc[:2], att[:2], None, None, 3, 1, xe_mode=True)
    out["cfg1_beam3xe_tokens"], out["cfg1_beam3xe_lens"], out["cfg1_beam3xe_scores"] = tk, ln, sc
    # teacher-forced xe / seq2seq log-probs at the target ids
    caps = syn.synthetic_captions(B, V, T + 1, seed=2)
    with torch.no_grad():
        lpx = m(fc, att, cpts, caps, labels, mode="xe")  # [B,T,V]
        out["cfg1_xe_lp_target"] = lpx.gather(2, caps[:, 1:].unsqueeze(2)).squeeze(2).numpy()
        out["cfg1_xe_lp_max"] = lpx.max(2).values.numpy()
        out["cfg1_xe_argmax"] = lpx.argmax(2).numpy()
        lps = m(caps, cpts, sentis, labels, mode="seq2seq")
        out["cfg1_s2s_lp_target"] = lps.gather(2, caps[:, 1:].unsqueeze(2)).squeeze(2).numpy()
        out["cfg1_s2s_argmax"] = lps.argmax(2).numpy()
    # one forward_step from a non-zero state (rl mode), top-8 per row
    g = torch.Generator().manual_seed(7)
    h0 = torch.randn(2, B, 512, generator=g) * 0.3
    c0 = torch.randn(2, B, 512, generator=g) * 0.3
    it = torch.randint(0, V, (B,), generator=g)
    with torch.no_grad():
        fcE = m.fc_embed(fc)
        a = m.att_embed(att.view(B, -1, 2048))
        pa = m.att2att(a)
        sw = m.word_embed(torch.cat([sentis.new_zeros(B, 1), sentis], 1))
        psw = m.senti2att(sw)
        sl = m.senti_label_embed(labels)
        lp1, (h1, c1) = m.forward_step(it, (h0, c0), fcE, a, pa, sw, psw, sl)
        m.attention._reset_weights()
    top = lp1.topk(8, dim=1)
    out["step_top_vals"], out["step_top_idx"] = top.values.numpy(), top.indices.numpy()
    out["step_h"], out["step_c"] = h1.numpy(), c1.numpy()

    # ---- EOS-heavy: V=64, B=64 ----------------------------------------------------------
    V2, B2 = 64, 64
    m2 = build_ref(V2, 5, eos_heavy=True)
    fc2, att2, cpts2, sentis2, labels2 = syn.synthetic_inputs(B2, V2, seed=11)
    with torch.no_grad():
        seq, lp, mask = m2(fc2, att2, cpts2, sentis2, labels2, T, 1, mode="rl")
    out["eos_greedy_seq"], out["eos_greedy_lp"], out["eos_greedy_mask"] = seq.numpy(), lp.numpy(), mask.numpy()
    for K, cons in ((3, 1), (5, 1), (3, 0)):
        tk, ln, sc = ref_beam(m2, fc2[:24], att2[:24], sentis2, labels2, K, cons)
        out[f"eos_beam{K}c{cons}_tokens"], out[f"eos_beam{K}c{cons}_lens"], out[f"eos_beam{K}c{cons}_scores"] = tk, ln, sc
    print("eos greedy lengths:", sorted(set(mask.sum(1).int().tolist())))
    print("eos beam3 top-1 lengths:", sorted(set(out["eos_beam3c1_lens"][:, 0].tolist())))
    np.savez_compressed(os.path.join(HERE, "decode_golden.npz"), **out)
    print("cfg1 greedy row0:", out["cfg1_greedy_seq"][0], out["cfg1_greedy_lp"][0, :3])
    print("cfg1 beam3 img0 scores:", out["cfg1_beam3_scores"][0])


def cider_goldens():
    V, N = 1000, 96
    refs = syn.synthetic_references(N, V, 5, seed=3)
    fns = ["img%d" % i for i in range(N)]
    split = {"train": {fn: refs[i] for i, fn in enumerate(fns[:64])},
             "val": {fn: refs[64 + i] for i, fn in enumerate(fns[64:])}}
    scorer = get_ciderd_scorer(split, 1, 2)
    g = torch.Generator().manual_seed(9)
    B = 64
    sample = torch.zeros(B, T, dtype=torch.long)
    greedy = torch.zeros(B, T, dtype=torch.long)
    for i in range(B):
        for dst, j in ((sample, 0), (greedy, 1)):
            r = refs[i][(i + j) % 5][1:-1]  # a reference, perturbed
            ids = list(r)
            for q in range(len(ids)):
                if float(torch.rand(1, generator=g)) < 0.3:
                    ids[q] = int(torch.randint(0, 60, (1,), generator=g))
            ids = ids[: int(torch.randint(0, len(ids) + 1, (1,), generator=g))] if i % 7 == 0 else ids
            ids = (ids + [2])[:T]
            dst[i, :len(ids)] = torch.tensor(ids)
    sample[5] = 0
    sample[5, 0] = 2  # EOS-only hypothesis
    greedy[6] = torch.randint(4, V, (T,), generator=g)  # no EOS at all, 16 words
    gt = {fn: refs[i] for i, fn in enumerate(fns[:B])}
    rewards = get_self_critical_reward(sample, greedy, fns[:B], gt, 1, 2, scorer)
    # raw scores for sample and greedy separately
    from self_critical.utils import _array_to_str
    res = [{"image_id": fns[i], "caption": [_array_to_str(sample[i].numpy(), 1, 2)]} for i in range(B)]
    res += [{"image_id": fns[i], "caption": [_array_to_str(greedy[i].numpy(), 1, 2)]} for i in range(B)]
    gts = {fns[i]: [_array_to_str(c, 1, 2) for c in refs[i]] for i in range(B)}
    _, scores = scorer.compute_score(gts, res)
    # n-gram count dump for the first 8 sample strings: sorted (ngram tuple as ints, count)
    dump = []
    for i in range(8):
        cnt = precook(_array_to_str(sample[i].numpy(), 1, 2))
        rows = sorted((tuple(int(w) for w in k), v) for k, v in cnt.items())
        flat = []
        for k, v in rows:
            flat.append(list(k) + [-1] * (4 - len(k)) + [v])
        dump.append(np.array(flat, dtype=np.int64))
    df_items = sorted((tuple(int(w) for w in k), v) for k, v in scorer.cider_scorer.document_frequency.items()
                      if v > 0)
    df_arr = np.array([list(k) + [-1] * (4 - len(k)) + [int(v)] for k, v in df_items], dtype=np.int64)
    out = {"sample": sample.numpy(), "greedy": greedy.numpy(), "rewards": rewards, "scores": scores,
           "ref_len": np.array(float(scorer.cider_scorer.ref_len)), "df": df_arr}
    for i, d in enumerate(dump):
        out[f"ngrams_{i}"] = d
    np.savez_compressed(os.path.join(HERE, "cider_golden.npz"), **out)
    print("cider scores head:", scores[:6], "EOS-only:", scores[5], "n df entries:", len(df_arr))


def senti_goldens():
    """The reference's SentimentDetector (models/sentiment_detector.py) on synthetic region features."""
    from models.sentiment_detector import SentimentDetector as RefDetector
    settings = dict(syn.DEFAULT_SETTINGS, sentiment_convs_num=2, sentiment_fcs_num=2)
    m = RefDetector(syn.SENTIMENT_CATEGORIES, settings)
    m.load_state_dict(syn.senti_detector_state_dict(0))
    m.eval()
    att = syn.senti_detector_inputs(6)
    with torch.no_grad():
        output, maps = m(att)
        labels, _, names, scores = m.sample(att, 0.7)
        m2 = RefDetector(["neutral", "positive", "negative"], settings)  # neutral = class 0: the threshold changes labels
        m2.load_state_dict(syn.senti_detector_state_dict(0))
        labels0, _, _, _ = m2.sample(att, 0.7)
    out = {"output": output.numpy(), "maps": maps.numpy(), "labels": labels.numpy(), "scores": scores.numpy(),
           "labels_neutral_first": labels0.numpy(), "checksum_att": np.array(checksum(att))}
    np.savez_compressed(os.path.join(HERE, "senti_golden.npz"), **out)
    print("senti scores:", scores.numpy(), "labels:", labels.numpy(), names)


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


def loader_goldens():
    """The reference's seven collate functions (dataloader.py:9-151) on synthetic dataset items. The module imports h5py
    at the top (absent here; only its Dataset classes use it), so an empty stand-in module is registered first."""
    import json
    import random
    import types
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    import dataloader as ref_dl
    it = syn.loader_items()
    fc, att = it["fc"].numpy(), it["att"].numpy()
    row = {fn: i for i, fn in enumerate(it["names"])}
    items = {
        "caption": [(fn, fc[row[fn]], att[row[fn]], caps, it["concepts"][fn]) for fn, caps in it["captions"].items()],
        "rl_fact": [(fn, caps, fc[row[fn]], att[row[fn]], it["concepts"][fn], it["sentiments"][fn]) for fn, caps in it["captions"].items()],
        "rl_senti": [(fn, fc[row[fn]], att[row[fn]], it["concepts"][fn], it["sentiments"][fn], lab) for fn, lab in it["labels"]],
        "senti_image": [(fn, att[row[fn]], lab) for fn, lab in it["labels"]],
        "concept": [(fn, fc[row[fn]], np.eye(1, 20, k=row[fn], dtype=np.int16)[0]) for fn in it["names"]],
        "senti_corpus_with_sentis": list(it["corpus"]),
        "senti_sents": [(lab, np.array(cap)) for cap, _, _, lab in it["corpus"]],
    }
    out = {}
    for name, batch in items.items():
        random.seed(5)  # rl_fact draws one caption per image with random.sample
        fn = ref_dl.create_collate_fn(name, pad_index=0, max_seq_len=17, num_concepts=5, num_sentiments=10)
        out[name] = jsonable(fn(list(batch)))
    with open(os.path.join(HERE, "loader_golden.json"), "w") as f:
        json.dump(out, f)
    print("loader goldens:", {k: len(json.dumps(v)) for k, v in out.items()})


def detector_goldens():
    """The reference's Detector (models/decoder.py:21-192) with its OWN SentimentDetector and SentenceSentimentClassifier on
    in-memory synthetic batches, ``training=False`` (dropout off, no scheduled sampling, no optimizer step). The sampled
    pass draws through torch.multinomial, whose stream no other implementation can reproduce: it is patched HERE (the
    reference files stay untouched) to Gumbel-max over injected noise, argmax(log p + g_t) — the form the B200 path's
    forward_rl(noise=...) consumes — so that sampled ids, both rewards and the REINFORCE loss become comparable."""
    from models.decoder import Detector as RefDetector
    V, B, Tm = 64, 6, 12  # the EOS-heavy recipe: peaked word distributions, captions of varied length
    settings = dict(syn.DEFAULT_SETTINGS, sentiment_convs_num=2, sentiment_fcs_num=2)
    torch.manual_seed(0)
    # neutral FIRST: with it last every synthetic image is labelled neutral (see senti_goldens); this order yields mixed
    # labels. The classifier seed is the first whose reward is non-zero on these captions (recorded in the golden).
    cats = ["neutral", "positive", "negative"]
    d = RefDetector(syn.make_vocab(V), Tm, cats, {"cap_lr": 4e-4}, settings)
    d.captioner.load_state_dict(syn.synthetic_state_dict(V, 0, eos_heavy=True))
    d.senti_detector.load_state_dict(syn.senti_detector_state_dict(0))
    cls_seed = int(os.environ.get("ISC_GOLDEN_CLS_SEED", "3"))
    d.sent_senti_cls.load_state_dict(syn.sent_cls_state_dict(V, cls_seed))
    fact, senti, scs, gts = syn.detector_batches(B, V, Tm)
    d.set_ciderd_scorer({"train": gts})
    noise = syn.gumbel_noise(Tm, B, V, seed=21)
    state = {"t": 0}
    real_multinomial = torch.multinomial

    def gumbel_max(probs, n, *a, **k):
        t = state["t"]
        state["t"] += 1
        return (probs.log() + noise[t]).argmax(dim=1, keepdim=True)

    calls = []
    real_forward = d.captioner.forward

    def recording_forward(*args, **kwargs):
        out = real_forward(*args, **kwargs)
        if kwargs.get("mode") == "rl":
            calls.append([o.detach().clone() for o in out])
        return out

    d.captioner.forward = recording_forward
    out = {"cls_seed": np.array(cls_seed), "categories": np.array(cats)}
    torch.multinomial = gumbel_max
    try:
        for tag, data, dtype in (("fact", (fact, scs), "fact"), ("senti", (senti, scs), "senti")):
            state["t"] = 0
            del calls[:]
            with torch.no_grad():  # evaluation call (train_rl.py:258 runs it under no_grad as well)
                losses = d(data, dtype, training=False)
            for k, v in losses.items():
                out["%s_%s" % (tag, k)] = np.array(float(v))
            (s_seq, s_lp, s_mask), (g_seq, g_lp, g_mask) = calls[0], calls[1]
            out[tag + "_sample_seq"] = s_seq.numpy()
            out[tag + "_sample_lp"] = s_lp.numpy()
            out[tag + "_sample_mask"] = s_mask.numpy()
            out[tag + "_greedy_seq"] = g_seq.numpy()
            print(tag, "sampled:", s_seq[:3].tolist(), "greedy:", g_seq[:3].tolist())
            print("detector", tag, {k: float(v) for k, v in losses.items()})
    finally:
        torch.multinomial = real_multinomial
        d.captioner.forward = real_forward
    fc, att, cpts, sentis, labels = syn.synthetic_inputs(2, V, seed=9)
    caps, sentiments = [], []
    for i in range(2):
        c, s_ = d.sample(fc[i], att[i], sentis[i], beam_size=3)
        caps.append(c)
        sentiments.append(s_)
    out["sample_captions"] = np.array(caps)
    out["sample_sentiments"] = np.array(sentiments)
    np.savez_compressed(os.path.join(HERE, "detector_golden.npz"), **out)
    print("detector sample:", caps[0][0], sentiments)


def sentcls_goldens():
    """The reference's SentenceSentimentClassifier (models/sent_senti_cls.py) on ragged synthetic captions."""
    from models.sent_senti_cls import SentenceSentimentClassifier as RefCls
    V, B = 1000, 12
    m = RefCls(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS))
    m.load_state_dict(syn.sent_cls_state_dict(V, 0))
    m.eval()
    seqs, lengths = syn.sent_cls_inputs(B, V)
    with torch.no_grad():
        pred, weights = m(seqs, lengths)
        result, names, _ = m.sample(seqs, lengths)
    out = {"pred": pred.numpy(), "weights": weights.numpy(), "result": np.array(result), "lengths": np.array(lengths),
           "checksum_seqs": np.array(checksum(seqs))}
    np.savez_compressed(os.path.join(HERE, "sentcls_golden.npz"), **out)
    print("sentcls pred:", pred.numpy()[:3], "result:", result, "lengths:", lengths)


if __name__ == "__main__":
    torch.set_num_threads(8)
    if "senti" in sys.argv:
        senti_goldens()
    elif "sentcls" in sys.argv:
        sentcls_goldens()
    elif "loader" in sys.argv:
        loader_goldens()
    elif "detector" in sys.argv:
        detector_goldens()
    else:
        detector_goldens()
        sentcls_goldens()
        loader_goldens()
        decode_goldens()
        cider_goldens()
        senti_goldens()
