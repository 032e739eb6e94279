"""Experiment: one beam-3 call over B images vs the same images as S independent sub-batches decoded concurrently on S
streams inside ONE captured graph (tail CTAs of one sub-batch's kernels fill in with the other's; HBM-bound attention
of one overlaps tensor-bound GEMMs of the other). Usage: python profiles/dual_stream_bench.py [B] [S]"""
import sys

import torch

from insenticap_model_b200 import synthetic as syn
from insenticap_model_b200.captioner import Captioner

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2
V = 10000
dev = torch.device("cuda:0")
sd = syn.synthetic_state_dict(V, 0)
models = []
for _ in range(S):
    m = Captioner(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS))
    m.load_state_dict(sd)
    models.append(m.to(dev).eval())
fc, att, cpts, sentis, labels = syn.synthetic_inputs(B, V, seed=1)
fc, att, sentis, labels = fc.to(dev), att.to(dev), sentis.to(dev), labels.to(dev)
att = att.reshape(B, 196, -1)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    m0 = models[0]
    m0.use_cuda_graph = True
    ref = [x.clone() for x in m0.beam_search(fc, att, sentis, labels, 3, 1, 16)]
    ms1 = timed(lambda: m0.beam_search(fc, att, sentis, labels, 3, 1, 16))
    print("single stream B=%d: %.3f ms, %.0f captions/s" % (B, ms1, B / ms1 * 1e3))

    per = B // S
    parts = [(fc[i * per:(i + 1) * per], att[i * per:(i + 1) * per], sentis[i * per:(i + 1) * per], labels[i * per:(i + 1) * per])
             for i in range(S)]
    for m, p in zip(models, parts):  # eager warm-up sizes the workspaces
        m.use_cuda_graph = False
        m.beam_search(*p, 3, 1, 16)
    torch.cuda.synchronize()
    side = [torch.cuda.Stream(dev) for _ in range(S - 1)]
    g = torch.cuda.CUDAGraph()
    outs = [None] * S
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        for s in side:
            s.wait_stream(cur)
        for i in range(1, S):
            with torch.cuda.stream(side[i - 1]):
                outs[i] = models[i]._beam_search_device(*parts[i], 3, 1, 16)
        outs[0] = models[0]._beam_search_device(*parts[0], 3, 1, 16)
        for s in side:
            cur.wait_stream(s)
    msS = timed(g.replay)
    tok = torch.cat([o[0] for o in outs])
    print("%d streams x B=%d: %.3f ms, %.0f captions/s, tokens identical: %s" % (S, per, msS, B / msS * 1e3, bool(torch.equal(tok, ref[0]))))
