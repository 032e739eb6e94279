"""GPU CIDEr-D reward vs the reference-generated goldens and the CPU oracle: n-gram counts
bit-exact, scores within 1e-5 (north star; actual agreement ~1e-13)."""
import numpy as np
import pytest
import torch

from insenticap_model_b200 import reward as R
from insenticap_model_b200 import synthetic as syn
from oracle import cider_oracle as C

pytestmark = pytest.mark.gpu
T = 16


def _setup():
    V, N = 1000, 96
    refs = syn.synthetic_references(N, V, 5, seed=3)
    fns = ["img%d" % i for i in range(N)]
    split = {"train": {fn: refs[i] for i, fn in enumerate(fns[:64])},
             "val": {fn: refs[64 + i] for i, fn in enumerate(fns[64:])}}
    scorer = R.get_ciderd_scorer(split, 1, 2, device="cuda")
    return refs, fns, scorer


def test_ngram_counts_bit_exact(golden_cider):
    refs, fns, scorer = _setup()
    for i in range(8):
        got = scorer.ngram_counts(torch.from_numpy(golden_cider["sample"][i]), 1, 2)
        want = {}
        for row in golden_cider[f"ngrams_{i}"]:
            want[C.pack_key([int(x) for x in row[:4] if x >= 0])] = int(row[4])
        assert got == want
    assert abs(scorer.ref_len - float(golden_cider["ref_len"])) < 1e-15


def test_scores_and_reward_match_reference(golden_cider):
    refs, fns, scorer = _setup()
    B = golden_cider["sample"].shape[0]
    sample = torch.from_numpy(golden_cider["sample"]).cuda()
    greedy = torch.from_numpy(golden_cider["greedy"]).cuda()
    gt = {fn: refs[i] for i, fn in enumerate(fns[:B])}
    scores = R.self_critical_scores(sample, greedy, fns[:B], gt, 1, 2, scorer).cpu().numpy()
    np.testing.assert_allclose(scores, golden_cider["scores"], rtol=0, atol=1e-10)
    rewards = R.get_self_critical_reward(sample, greedy, fns[:B], gt, 1, 2, scorer)
    assert isinstance(rewards, np.ndarray) and rewards.dtype == np.float64 and rewards.shape == (B, T)
    np.testing.assert_allclose(rewards, golden_cider["rewards"], rtol=0, atol=1e-10)
    assert scores[5] == 0.0  # EOS-only hypothesis
    # registered-ground-truth fast path gives the same numbers; device tensor out
    scorer.register_ground_truth(fns[:B], gt, 1, 2)
    r2 = R.get_self_critical_reward(sample, greedy, fns[:B], gt, 1, 2, scorer, as_tensor=True)
    assert r2.is_cuda and np.allclose(r2.cpu().numpy(), rewards, atol=0)


def test_string_api_compute_score(golden_cider):
    """CiderD.compute_score(gts, res) with the reference's space-joined id strings (ciderD.py:24-48)."""
    refs, fns, scorer = _setup()
    B = 16
    to_str = lambda ids: " ".join(str(w) for w in C.ids_to_words(ids, 1, 2))
    res = [{"image_id": fns[i], "caption": [to_str(golden_cider["sample"][i])]} for i in range(B)]
    res += [{"image_id": fns[i], "caption": [to_str(golden_cider["greedy"][i])]} for i in range(B)]
    gts = {fns[i]: [to_str(c) for c in refs[i]] for i in range(B)}
    mean, scores = scorer.compute_score(gts, res)
    want = np.concatenate([golden_cider["scores"][:B], golden_cider["scores"][64:64 + B]])
    np.testing.assert_allclose(scores, want, rtol=0, atol=1e-10)
    assert abs(mean - want.mean()) < 1e-10


def test_large_random_against_oracle():
    V, N = 3000, 512
    refs = syn.synthetic_references(N, V, 5, seed=13)
    orc = C.CiderOracle(refs, 1, 2)
    scorer = R.CiderD(refs=[[C.ids_to_words(c, 1, 2) for c in caps] for caps in refs], device="cuda")
    g = torch.Generator().manual_seed(17)
    hyps = torch.zeros(2 * N, T, dtype=torch.long)
    for i in range(2 * N):
        base = refs[i % N][int(torch.randint(0, 5, (1,), generator=g))][1:-1]
        ids = [w if float(torch.rand(1, generator=g)) > 0.25 else int(torch.randint(0, 40, (1,), generator=g)) for w in base]
        ids = (ids + [2])[:T]
        if i % 11 == 0:
            ids = ids[:-1] if len(ids) == T else ids  # some rows without EOS inside T
        hyps[i, :len(ids)] = torch.tensor(ids)
    refset = R.RefSet([[C.ids_to_words(c, 1, 2) for c in caps] for caps in refs], "cuda")
    img = torch.arange(2 * N, dtype=torch.int32) % N
    got = scorer.score_ids(hyps.cuda(), img, refset, 1, 2).cpu().numpy()
    want = np.array([orc.score(hyps[i].tolist(), refs[i % N]) for i in range(2 * N)])
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-9)
    assert got.max() > 1.0 and (got == 0).sum() < N  # a meaningful spread of scores
