"""GPU parity of the sentence sentiment classifier (SentenceSentimentClassifier.forward / .sample on libisc_b200.so: the
LSTM on the fused gate GEMM, the excitation / classifier layers on the tensor-core GEMM) against the CPU oracle and the
reference-generated golden (tests/golden/sentcls_golden.npz). Logits and word weights within 1e-4, classes identical."""
import os

import numpy as np
import pytest
import torch

from insenticap_model_b200 import synthetic as syn
from insenticap_model_b200.sent_senti_cls import SentenceSentimentClassifier
from oracle import sentcls_oracle as CO

pytestmark = pytest.mark.gpu


def _model(V):
    m = SentenceSentimentClassifier(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS))
    m.load_state_dict(syn.sent_cls_state_dict(V, 0))
    return m.cuda().eval()


def test_sentcls_matches_reference_golden():
    gd = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sentcls_golden.npz"))
    V, B = 1000, 12
    seqs, lengths = syn.sent_cls_inputs(B, V)
    m = _model(V)
    pred, weights = m(seqs.cuda(), lengths)
    result, names, w2 = m.sample(seqs.cuda(), lengths)
    torch.cuda.synchronize()
    assert tuple(weights.shape) == gd["weights"].shape
    np.testing.assert_allclose(pred.cpu().numpy(), gd["pred"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(weights.cpu().numpy(), gd["weights"], rtol=1e-4, atol=1e-5)
    assert result == gd["result"].tolist() and names == [syn.SENTIMENT_CATEGORIES[r] for r in result]
    assert torch.equal(w2, weights)


@pytest.mark.parametrize("B,T", [(1, 1), (200, 20), (1024, 17)])
def test_sentcls_ragged_batches_match_oracle(B, T):
    """Ragged lengths in 1..T (one caption of length 1, one full), batch sizes either side of a 128-row tile."""
    V = 1000
    seqs, lengths = syn.sent_cls_inputs(B, V, max_len=T, seed=40 + B)
    m = _model(V)
    pred, weights = m(seqs.cuda(), torch.tensor(lengths))
    torch.cuda.synchronize()
    with torch.no_grad():
        want, ww = CO.forward(syn.sent_cls_state_dict(V, 0), seqs, lengths)
    np.testing.assert_allclose(pred.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(weights.cpu().numpy(), ww.numpy(), rtol=1e-4, atol=1e-5)
    valid = torch.arange(max(lengths))[None, :] < torch.tensor(lengths)[:, None]
    assert float(weights.cpu()[~valid].abs().sum()) == 0.0  # nothing past a caption's length
    # a caption's result does not depend on the batch it is in, nor on padding columns beyond max(lengths)
    if B >= 8:
        sub = slice(3, 8)
        wide = torch.cat([seqs[sub], torch.full((5, 4), 7, dtype=seqs.dtype)], dim=1)
        p2, w2 = m(wide.cuda(), lengths[sub])
        t2 = max(lengths[sub])
        np.testing.assert_allclose(p2.cpu().numpy(), pred[sub].cpu().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(w2.cpu().numpy(), weights[sub, :t2].cpu().numpy(), rtol=1e-5, atol=1e-6)


def test_sentcls_errors():
    V = 100
    seqs, lengths = syn.sent_cls_inputs(4, V, max_len=6)
    m = _model(V)
    with pytest.raises(RuntimeError, match="no CPU"):
        m(seqs, lengths)
    with pytest.raises(RuntimeError, match=">= 1"):
        m(seqs.cuda(), [0, 1, 2, 3])
    with pytest.raises(RuntimeError, match="width"):
        m(seqs.cuda(), [7, 1, 2, 3])
    m.train()
    with pytest.raises(NotImplementedError):
        m(seqs.cuda(), lengths)
