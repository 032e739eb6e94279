"""Profiling driver: ONE beam-3 decode call (B=1024, V=10000) after packing weights. Used under ncu:
   launch list : ncu --metrics gpu__time_duration.sum --clock-control none --csv ...
   full capture: ncu --set full -k regex:'gemm_tc|attention|beam_select' -s 43 -c 7 ...   (decode step t=1)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from insenticap_model_b200 import synthetic as syn  # noqa: E402
from insenticap_model_b200.captioner import Captioner  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 1
V = 10000
m = Captioner(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS), precision=precision)
m.load_state_dict(syn.synthetic_state_dict(V, 0))
m = m.cuda().eval()
g = torch.Generator(device="cuda").manual_seed(1)
fc = torch.rand(B, 2048, device="cuda", generator=g)
att = torch.rand(B, 14, 14, 2048, device="cuda", generator=g)
sentis = torch.randint(4, V, (B, 10), device="cuda", generator=g)
labels = (torch.arange(B, device="cuda") % 3).long()
for _ in range(calls):
    tk, sc, ln = m.beam_search(fc, att, sentis, labels, beam_size=3, max_seq_len=16)
torch.cuda.synchronize()
print("ok", tk.shape, float(sc[0, 0]))
