"""Host-side logic that needs no GPU: drop-in surface, state_dict compatibility, loud failure on CPU,
reference-set packing, criteria."""
import numpy as np
import pytest
import torch

from insenticap_model_b200 import synthetic as syn
from insenticap_model_b200.captioner import Captioner, XECriterion
from insenticap_model_b200 import reward as R


def _cap(V=50):
    return Captioner(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS))


def test_state_dict_keys_and_shapes_match_reference_layout():
    V = 50
    m = _cap(V)
    want = [(n, s) for n, s, _ in syn.param_specs(V)]
    got = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert got == want  # same 40 names, same order, same shapes as the reference's Captioner
    m.load_state_dict(syn.synthetic_state_dict(V, 0))  # loads verbatim
    assert (m.pad_id, m.sos_id, m.eos_id, m.unk_id, m.neu_idx, m.vocab_size) == (0, 1, 2, 3, 2, V)
    h, c = m.init_hidden(4)
    assert h.shape == (2, 4, 512) and float(h.abs().sum()) == 0.0 and c.shape == (2, 4, 512)


def test_vocab_without_sos_uses_pad_for_sos_and_eos():
    m = Captioner(["<PAD>", "<UNK>", "a", "b", "c", "d", "e", "f"], ["neutral", "positive", "negative"],
                  dict(syn.DEFAULT_SETTINGS))
    assert m.sos_id == m.pad_id == m.eos_id == 0 and m.neu_idx == 0


def test_no_cpu_fallback_and_mode_dispatch():
    m = _cap().eval()
    fc, att, cpts, sentis, labels = syn.synthetic_inputs(2, 50)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.forward_rl(fc, att, cpts, sentis, labels, 16, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.sample(fc[0], att[0])
    with pytest.raises(AttributeError):
        m(fc, mode="nonexistent")  # reference: getattr(self, 'forward_' + mode), captioner.py:192
    m.train()  # the training path (tape + hand-written backward) is CUDA-only as well
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.forward_xe(fc, att, cpts, syn.synthetic_captions(2, 50), labels)
    with pytest.raises(RuntimeError, match="no CPU fallback"):  # scheduled sampling: same path, same rule
        m._teacher_forced_train(0, fc, att, cpts, None, labels, syn.synthetic_captions(2, 50), 0.25)
    with pytest.raises(ValueError):
        Captioner(syn.make_vocab(50), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS, rnn_hid_dim=256))


def test_xe_criterion_matches_formula():
    torch.manual_seed(0)
    pred = torch.log_softmax(torch.randn(3, 5, 7), -1)
    target = torch.randint(0, 7, (3, 5))
    lengths = [5, 3, 1]
    loss = XECriterion()(pred, target, lengths)
    want = 0.0
    for i, l in enumerate(lengths):
        for t in range(l):
            want -= float(pred[i, t, target[i, t]])
    assert abs(float(loss) - want / sum(lengths)) < 1e-6
    opt, xe, mse = _cap().get_optim_criterion(1e-3)
    assert isinstance(opt, torch.optim.Adam) and isinstance(xe, XECriterion) and isinstance(mse, torch.nn.MSELoss)


def test_reward_host_packing():
    assert R.ids_to_words([1, 5, 6, 2, 9], 1, 2) == [5, 6, 2]
    assert R.ids_to_words([5, 6], 1, 2) == [5, 6, 2]
    assert R.ids_to_words([2, 0, 0], 1, 2) == [2]
    rs = R.RefSet([[[5, 6, 2], [7, 2]], [[8, 9, 10, 2]]], "cpu")
    assert rs.offsets.tolist() == [0, 2, 3] and rs.lens.tolist() == [3, 2, 4] and rs.ld == 4
    assert rs.tokens.tolist() == [[5, 6, 2, 0], [7, 2, 0, 0], [8, 9, 10, 2]]
    with pytest.raises(ValueError):
        R.RefSet([[list(range(40))]], "cpu")
    with pytest.raises(ValueError):
        R.RefSet([[[70000, 2]]], "cpu")
    with pytest.raises(RuntimeError, match="GPU only"):
        R.CiderD(device="cpu")
    with pytest.raises(ValueError):
        R.CiderD(n=5, device="cuda")
    crit = R.RewardCriterion()
    lp = torch.tensor([[-1.0, -2.0], [-3.0, -4.0]])
    mk = torch.tensor([[1.0, 1.0], [1.0, 0.0]])
    rw = torch.tensor([[0.5, 0.5], [2.0, 2.0]])
    assert abs(float(crit(lp, mk, rw)) - (0.5 + 1.0 + 6.0) / 3) < 1e-6


def test_fused_adam_state_dict_and_param_groups_roundtrip():
    """FusedClampAdam offers the optimizer surface the reference scripts use: param_groups[...]['lr'] writes through
    (train_xe.py:130-133) and state_dict()/load_state_dict() carry moments, step count and hyper-parameters (train_xe.py:53, :245)."""
    import torch
    from insenticap_model_b200 import train as TR
    lin = torch.nn.Linear(4, 3)
    lin._packed_key = None
    opt = TR.FusedClampAdam(lin, lr=1e-3, grad_clip=0.1)
    for g in opt.param_groups:
        g["lr"] = 5e-4
    assert opt.lr == 5e-4 and opt.param_groups[0]["lr"] == 5e-4 and len(opt.param_groups[0]["params"]) == 2
    opt.exp_avg.fill_(0.25)
    opt.exp_avg_sq.fill_(0.5)
    opt.steps = 7
    sd = opt.state_dict()
    lin2 = torch.nn.Linear(4, 3)
    lin2._packed_key = None
    opt2 = TR.FusedClampAdam(lin2, lr=1.0)
    opt2.load_state_dict(sd)
    assert opt2.steps == 7 and opt2.lr == 5e-4 and opt2.grad_clip == 0.1
    assert torch.equal(opt2.exp_avg, opt.exp_avg) and torch.equal(opt2.exp_avg_sq, opt.exp_avg_sq)
    import pytest
    with pytest.raises(ValueError):
        TR.FusedClampAdam(torch.nn.Linear(2, 2), lr=1.0).load_state_dict(sd)
