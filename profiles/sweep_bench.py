"""BASELINE configs[4]: throughput sweep of Captioner.beam_search over beam sizes and batch sizes on one B200 (V = 10000,
16 tokens, bf16x3, CUDA-graph replay, device-resident inputs). Usage: python profiles/sweep_bench.py [precision]
Prints one JSON line per (beam, batch) and a markdown table at the end."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from insenticap_model_b200 import synthetic as syn  # noqa: E402
from insenticap_model_b200.captioner import Captioner  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
V, T = 10000, 16
m = Captioner(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS), precision=precision)
m.load_state_dict(syn.synthetic_state_dict(V, 0))
m = m.cuda().eval()
m.use_cuda_graph = True
rows = []
for B in (64, 256, 1024, 4096, 8192):
    g = torch.Generator(device="cuda").manual_seed(B)
    fc = torch.rand(B, 2048, device="cuda", generator=g)
    att = torch.rand(B, 14, 14, 2048, device="cuda", generator=g)
    sentis = torch.randint(4, V, (B, 10), device="cuda", generator=g)
    labels = (torch.arange(B, device="cuda") % 3).long()
    for K in (1, 3, 5):
        m._graphs.clear()
        step = lambda: m.beam_search(fc, att, sentis, labels, beam_size=K, max_seq_len=T)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        n = max(3, min(20, int(2e4 / (B * K) ** 0.9) + 3))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        r = {"beam": K, "batch": B, "ms_per_call": ms, "captions_per_s": B / (ms * 1e-3), "precision": precision}
        rows.append(r)
        print(json.dumps(r), flush=True)
    del fc, att
    m._ws.clear()
    torch.cuda.empty_cache()
print("\n| batch | beam 1 | beam 3 | beam 5 |  (captions/s, %s)\n|---|---|---|---|" % precision)
for B in (64, 256, 1024, 4096, 8192):
    print("| %d | " % B + " | ".join("%.0f" % next(r["captions_per_s"] for r in rows if r["batch"] == B and r["beam"] == K)
                                      for K in (1, 3, 5)) + " |")
