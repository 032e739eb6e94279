"""GPU parity of the image sentiment detector (SentimentDetector.forward / .sample on libisc_b200.so: the two 3x3
convolutions as tcgen05 GEMMs over the zero-bordered 16x16 grid) against the CPU oracle and the reference-generated
golden (tests/golden/senti_golden.npz). Logits within 1e-4, labels identical (thresholded arg-max)."""
import os

import numpy as np
import pytest
import torch

from insenticap_model_b200 import synthetic as syn
from insenticap_model_b200.sentiment_detector import SentimentDetector
from oracle import senti_oracle as SO

pytestmark = pytest.mark.gpu
SETTINGS = dict(syn.DEFAULT_SETTINGS, sentiment_convs_num=2, sentiment_fcs_num=2)


def _model(cats=None):
    m = SentimentDetector(cats or syn.SENTIMENT_CATEGORIES, SETTINGS)
    m.load_state_dict(syn.senti_detector_state_dict(0))
    return m.cuda().eval()


def test_detector_matches_reference_golden_and_oracle():
    gd = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "senti_golden.npz"))
    att = syn.senti_detector_inputs(6)
    m = _model()
    out, maps = m(att.cuda())
    labels, maps2, names, scores = m.sample(att.cuda(), 0.7)
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), gd["output"], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(maps.cpu().numpy(), gd["maps"], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(scores.cpu().numpy(), gd["scores"], rtol=1e-4, atol=1e-4)
    assert np.array_equal(labels.cpu().numpy(), gd["labels"]) and names == ["neutral"] * 6
    labels0 = _model(["neutral", "positive", "negative"]).sample(att.cuda(), 0.7)[0]
    assert np.array_equal(labels0.cpu().numpy(), gd["labels_neutral_first"])


def test_detector_batch_larger_than_a_chunk_and_errors():
    """B = 80 spans two 74-image passes; results must not depend on the batch an image is in (bit for bit)."""
    att = syn.senti_detector_inputs(80, seed=23)
    m = _model()
    out, maps = m(att.cuda())
    out6, maps6 = m(att[70:76].cuda())
    torch.cuda.synchronize()
    assert torch.equal(out[70:76], out6) and torch.equal(maps[70:76], maps6)
    sd = syn.senti_detector_state_dict(0)
    with torch.no_grad():
        want, wmaps = SO.forward(sd, att[:3])
    np.testing.assert_allclose(out[:3].cpu().numpy(), want.numpy(), rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(maps[:3].cpu().numpy(), wmaps.numpy(), rtol=1e-4, atol=2e-4)
    with pytest.raises(RuntimeError, match="no CPU"):
        m(att[:1])
    m.train()
    with pytest.raises(NotImplementedError):
        m(att[:1].cuda())
