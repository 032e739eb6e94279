"""Pin the CPU oracle against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py, run in the authoring container). CPU only."""
import numpy as np
import torch

from insenticap_model_b200 import synthetic as syn
from oracle import captioner_oracle as O
from oracle import cider_oracle as C

T = 16


def _cfg1():
    V, B = 10000, 8
    p = syn.synthetic_state_dict(V, 0)
    inp = syn.synthetic_inputs(B, V, seed=1)
    return V, B, p, inp


def test_synthetic_tensors_reproduce(golden_decode):
    V, B, p, (fc, att, cpts, sentis, labels) = _cfg1()
    got = np.array([float(t.double().sum()) for t in (fc, att, cpts, sentis)])
    np.testing.assert_allclose(got, golden_decode["cfg1_checksum_inputs"], rtol=0, atol=0)
    gotw = np.array([float(v.double().sum()) for v in p.values()])
    np.testing.assert_allclose(gotw, golden_decode["cfg1_checksum_weights"], rtol=0, atol=0)


def test_greedy_cfg1(golden_decode):
    V, B, p, (fc, att, cpts, sentis, labels) = _cfg1()
    with torch.no_grad():
        f = O.prologue(p, fc, att, cpts, sentis, labels)
        seq, lp, mask = O.decode_greedy(p, f, B, T)
    assert np.array_equal(seq.numpy(), golden_decode["cfg1_greedy_seq"])
    np.testing.assert_allclose(lp.numpy(), golden_decode["cfg1_greedy_lp"], rtol=1e-5, atol=1e-5)
    assert np.array_equal(mask.numpy(), golden_decode["cfg1_greedy_mask"])
    np.testing.assert_allclose(f["fc_embedded"].numpy(), golden_decode["cfg1_fc_embedded"], atol=1e-5)
    np.testing.assert_allclose(f["cpt_feats"].numpy(), golden_decode["cfg1_cpt_feats"], atol=1e-5)


def test_step_from_nonzero_state(golden_decode):
    V, B, p, (fc, att, cpts, sentis, labels) = _cfg1()
    g = torch.Generator().manual_seed(7)
    h0 = torch.randn(2, B, 512, generator=g) * 0.3
    c0 = torch.randn(2, B, 512, generator=g) * 0.3
    it = torch.randint(0, V, (B,), generator=g)
    with torch.no_grad():
        f = O.prologue(p, fc, att, cpts, sentis, labels)
        lp, (h1, c1), (cw, sw, gw) = O.step(p, it, (h0, c0), f, want_weights=True)
    top = lp.topk(8, dim=1)
    assert np.array_equal(top.indices.numpy(), golden_decode["step_top_idx"])
    np.testing.assert_allclose(top.values.numpy(), golden_decode["step_top_vals"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(h1.numpy(), golden_decode["step_h"], atol=2e-6)
    np.testing.assert_allclose(c1.numpy(), golden_decode["step_c"], atol=2e-6)
    assert cw.shape == (B, 196) and sw.shape == (B, 11) and gw.shape == (B, 1)


def test_beam3_cfg1(golden_decode):
    V, B, p, (fc, att, cpts, sentis, labels) = _cfg1()
    with torch.no_grad():
        f = O.prologue(p, fc, att, None, sentis, labels)
        tk, sc, ln = O.beam_search(p, f, B, 3, 1, T)
    assert np.array_equal(tk.numpy(), golden_decode["cfg1_beam3_tokens"])
    assert np.array_equal(ln.numpy(), golden_decode["cfg1_beam3_lens"])
    # batch-K vs the reference's batch-1 GEMMs differ in fp32 rounding only
    np.testing.assert_allclose(sc.numpy(), golden_decode["cfg1_beam3_scores"], rtol=0, atol=2e-4)


def test_beam3_xe_mode_and_per_image(golden_decode):
    V, B, p, (fc, att, cpts, sentis, labels) = _cfg1()
    with torch.no_grad():
        f = O.prologue(p, fc[:2], att[:2])
        tk, sc, ln = O.beam_search(p, f, 2, 3, 1, T)
        assert np.array_equal(tk.numpy(), golden_decode["cfg1_beam3xe_tokens"])
        np.testing.assert_allclose(sc.numpy(), golden_decode["cfg1_beam3xe_scores"], atol=2e-4)
        f1 = O.prologue(p, fc[:1], att[:1])
        words, scores = O.beam_search_per_image(p, f1, 3, 1, T)
    for k in range(3):
        n = int(golden_decode["cfg1_beam3xe_lens"][0, k])
        assert words[k] == golden_decode["cfg1_beam3xe_tokens"][0, k, :n].tolist()
    np.testing.assert_allclose(scores, golden_decode["cfg1_beam3xe_scores"][0], atol=1e-5)


def test_teacher_forced_xe_and_seq2seq(golden_decode):
    V, B, p, (fc, att, cpts, sentis, labels) = _cfg1()
    caps = syn.synthetic_captions(B, V, T + 1, seed=2)
    with torch.no_grad():
        f = O.prologue(p, fc, att, cpts, None, labels)
        lp = O.teacher_forced(p, f, caps)
        tgt = lp.gather(2, caps[:, 1:].unsqueeze(2)).squeeze(2)
        np.testing.assert_allclose(tgt.numpy(), golden_decode["cfg1_xe_lp_target"], rtol=1e-5, atol=1e-5)
        assert np.array_equal(lp.argmax(2).numpy(), golden_decode["cfg1_xe_argmax"])
        f2 = O.prologue(p, None, None, cpts, sentis, labels, seq2seq=True)
        lp2 = O.teacher_forced(p, f2, caps)
        tgt2 = lp2.gather(2, caps[:, 1:].unsqueeze(2)).squeeze(2)
        np.testing.assert_allclose(tgt2.numpy(), golden_decode["cfg1_s2s_lp_target"], rtol=1e-5, atol=1e-5)
        assert np.array_equal(lp2.argmax(2).numpy(), golden_decode["cfg1_s2s_argmax"])


def test_eos_heavy_greedy_and_beams(golden_decode):
    V, B = 64, 64
    p = syn.synthetic_state_dict(V, 5, eos_heavy=True)
    fc, att, cpts, sentis, labels = syn.synthetic_inputs(B, V, seed=11)
    with torch.no_grad():
        f = O.prologue(p, fc, att, cpts, sentis, labels)
        seq, lp, mask = O.decode_greedy(p, f, B, T)
        assert np.array_equal(seq.numpy(), golden_decode["eos_greedy_seq"])
        assert np.array_equal(mask.numpy(), golden_decode["eos_greedy_mask"])
        np.testing.assert_allclose(lp.numpy(), golden_decode["eos_greedy_lp"], rtol=1e-4, atol=1e-4)
        assert len(set(mask.sum(1).int().tolist())) >= 5  # lengths really vary
        f24 = {k: (v[:24] if v is not None else None) for k, v in f.items()}
        for K, cons in ((3, 1), (5, 1), (3, 0)):
            tk, sc, ln = O.beam_search(p, f24, 24, K, cons, T)
            assert np.array_equal(tk.numpy(), golden_decode[f"eos_beam{K}c{cons}_tokens"]), (K, cons)
            assert np.array_equal(ln.numpy(), golden_decode[f"eos_beam{K}c{cons}_lens"])
            np.testing.assert_allclose(sc.numpy(), golden_decode[f"eos_beam{K}c{cons}_scores"], atol=5e-4)


# ------------------------------------------------------------------------------- CIDEr-D
def _cider_setup():
    V, N = 1000, 96
    refs = syn.synthetic_references(N, V, 5, seed=3)
    return refs, C.CiderOracle(refs, 1, 2)


def test_cider_ngram_counts_and_df(golden_cider):
    refs, orc = _cider_setup()
    for i in range(8):
        words = C.ids_to_words(golden_cider["sample"][i], 1, 2)
        cnt = C.ngram_counts(words)
        want = {}
        for row in golden_cider[f"ngrams_{i}"]:
            toks = [int(x) for x in row[:4] if x >= 0]
            want[C.pack_key(toks)] = int(row[4])
        assert cnt == want
    want_df = {C.pack_key([int(x) for x in row[:4] if x >= 0]): float(row[4]) for row in golden_cider["df"]}
    assert orc.df == want_df
    assert abs(orc.ref_len - float(golden_cider["ref_len"])) < 1e-15


def test_cider_scores_and_reward(golden_cider):
    refs, orc = _cider_setup()
    B = golden_cider["sample"].shape[0]
    s = [orc.score(golden_cider["sample"][i], refs[i]) for i in range(B)]
    g = [orc.score(golden_cider["greedy"][i], refs[i]) for i in range(B)]
    np.testing.assert_allclose(np.array(s + g), golden_cider["scores"], rtol=0, atol=1e-12)
    r = orc.self_critical_reward(golden_cider["sample"], golden_cider["greedy"], refs[:B])
    np.testing.assert_allclose(r, golden_cider["rewards"], rtol=0, atol=1e-12)
    assert golden_cider["scores"][5] == 0.0  # EOS-only hypothesis


def test_senti_detector_oracle_matches_reference_golden():
    """oracle/senti_oracle.py against the reference SentimentDetector's own outputs (tests/golden/senti_golden.npz)."""
    import os
    from oracle import senti_oracle as SO
    gd = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "senti_golden.npz"))
    sd = syn.senti_detector_state_dict(0)
    att = syn.senti_detector_inputs(6)
    assert abs(float(att.double().sum()) - float(gd["checksum_att"])) < 1e-6 * abs(float(gd["checksum_att"])) + 1e-6
    with torch.no_grad():
        out, maps = SO.forward(sd, att)
        labels, _, scores = SO.sample(sd, att, 0.7, 2)
        labels0, _, _ = SO.sample(sd, att, 0.7, 0)
    np.testing.assert_allclose(out.numpy(), gd["output"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(maps.numpy(), gd["maps"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(scores.numpy(), gd["scores"], rtol=1e-4, atol=1e-5)
    assert np.array_equal(labels.numpy(), gd["labels"]) and np.array_equal(labels0.numpy(), gd["labels_neutral_first"])
    assert len(set(gd["labels_neutral_first"].tolist())) > 1  # the threshold branch is exercised


def test_sentcls_oracle_matches_reference_golden():
    """oracle/sentcls_oracle.py against the reference SentenceSentimentClassifier's own outputs."""
    import os
    from oracle import sentcls_oracle as CO
    gd = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sentcls_golden.npz"))
    V, B = 1000, 12
    seqs, lengths = syn.sent_cls_inputs(B, V)
    assert abs(float(seqs.double().sum()) - float(gd["checksum_seqs"])) < 1e-6 and lengths == gd["lengths"].tolist()
    with torch.no_grad():
        pred, weights = CO.forward(syn.sent_cls_state_dict(V, 0), seqs, lengths)
    np.testing.assert_allclose(pred.numpy(), gd["pred"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(weights.numpy(), gd["weights"], rtol=1e-5, atol=1e-6)
    assert pred.argmax(-1).tolist() == gd["result"].tolist()


def test_port_matches_unmodified_reference_bytecode():
    """When oracle/_ref exists (built by oracle/build_ref.py from /root/reference), the oracle port and the UNMODIFIED
    reference `Captioner.sample` / `forward(mode='rl')` must agree on fresh inputs (not only on the committed goldens)."""
    import pytest
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ref = build_ref.import_reference()
    V, B = 1000, 3
    p = syn.synthetic_state_dict(V, 5)
    fc, att, cpts, sentis, labels = syn.synthetic_inputs(B, V, seed=11)
    m = ref.Captioner(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS))
    m.load_state_dict(p)
    m.eval()
    with torch.no_grad():
        seq_r, lp_r, _ = m(fc, att, cpts, sentis, labels, T, 1, mode="rl")
        f = O.prologue(p, fc, att, cpts, sentis, labels)
        seq, lp, _ = O.decode_greedy(p, f, B, T)
        assert np.array_equal(seq.numpy(), seq_r.numpy())
        np.testing.assert_allclose(lp.numpy(), lp_r.numpy(), rtol=1e-5, atol=1e-5)
        tk, sc, ln = O.beam_search(p, O.prologue(p, fc, att, None, sentis, labels), B, 3, 1, T)
        for i in range(B):
            caps, scores = m.sample(fc[i], att[i], sentis[i], labels[i:i + 1], beam_size=3, decoding_constraint=1, max_seq_len=T)
            for k in range(3):
                n = int(ln[i, k])
                words = [w for w in tk[i, k, :n].tolist() if w != 2]
                assert caps[k] == " ".join(m.idx2word[w] for w in words)
            np.testing.assert_allclose(scores, sc[i].numpy(), atol=2e-4)
