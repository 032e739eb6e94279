"""Shared helpers for the GPU parity tests."""
import torch

from insenticap_model_b200 import synthetic as syn
from insenticap_model_b200.captioner import Captioner

T = 16
_cache = {}


def model(V, seed, precision, eos_heavy=False):
    key = (V, seed, precision, eos_heavy)
    if key not in _cache:
        m = Captioner(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS), precision=precision)
        m.load_state_dict(syn.synthetic_state_dict(V, seed, eos_heavy=eos_heavy))
        _cache[key] = m.cuda().eval()
    return _cache[key]


def params(V, seed, eos_heavy=False):
    return syn.synthetic_state_dict(V, seed, eos_heavy=eos_heavy)


def to_cuda(*ts):
    return [t.cuda() if t is not None else None for t in ts]


def greedy_mismatch_report(seq, ref_seq, margins, tol):
    """Token mismatches are only acceptable where the oracle's top-2 margin is within `tol`
    (a numerical tie); after a tie-flip the rest of that row legitimately diverges."""
    seq, ref_seq = seq.cpu(), ref_seq.cpu()
    bad = []
    n_tie_rows = 0
    for b in range(seq.shape[0]):
        neq = (seq[b] != ref_seq[b]).nonzero().flatten()
        if len(neq) == 0:
            continue
        t0 = int(neq[0])
        if float(margins[b, t0]) <= tol:
            n_tie_rows += 1
        else:
            bad.append((b, t0, float(margins[b, t0])))
    return bad, n_tie_rows
