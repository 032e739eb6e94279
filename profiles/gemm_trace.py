"""Phase timeline of the fused-LSTM GEMM from a -DISC_GEMM_TRACE build of the library (per-CTA globaltimer stamps).
Build: nvcc ... -DISC_GEMM_TRACE -c csrc/gemm_tc.cu, link with the other objects -> ab/libT.so; then
  ISC_B200_LIB=$PWD/ab/libT.so python profiles/gemm_trace.py [images] [steps]
Runs a B x beam-3 call of `steps` decode steps and prints, for the LAST fused-LSTM GEMM launched (the language LSTM,
K = 1536, of the last step), the median / min / max over CTAs of each phase boundary in us."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from insenticap_model_b200 import _lib, synthetic as syn  # noqa: E402
from insenticap_model_b200.captioner import Captioner  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
# with a -DISC_TRACE_AF build the stamps come from the prologue's att_embed GEMM (last chunk) instead
V = 10000
lib = _lib.load()
raw = C.CDLL(os.environ["ISC_B200_LIB"])
m = Captioner(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS))
m.load_state_dict(syn.synthetic_state_dict(V, 0))
m = m.cuda().eval()
fc, att, cpts, sentis, labels = syn.synthetic_inputs(B, V, seed=1)
args = (fc.cuda(), att.cuda(), sentis.cuda(), labels.cuda())
with torch.no_grad():
    m.beam_search(*args, 3, 1, steps)
    torch.cuda.synchronize()
    trace = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    assert raw.isc_debug_gemm_trace(C.c_void_p(trace.data_ptr())) == 0
    m.beam_search(*args, 3, 1, steps)
    torch.cuda.synchronize()
    raw.isc_debug_gemm_trace(C.c_void_p(0))
t = trace.cpu().numpy().reshape(148, 16).astype(np.float64)
t0 = t[:, 0][t[:, 0] > 0].min()
std = "--std" in sys.argv  # plain-epilogue trace build: slots 3-6 are epilogue chunk stamps, not MMA / producer stamps
names = {0: "kernel entry", 1: "set-up done", 5: "stage 0 loads issued", 2: "first operand stage landed",
         12: "k-block 8: producer waits for a free stage", 13: "k-block 8: stage free", 6: "k-block 8: loads issued",
         14: "k-block 8: fp32 A tile in staging (converter)", 7: "k-block 8: stage landed (MMA thread)", 3: "tile0 MMAs issued", 4: "tile1 MMAs issued",
         8: "tile0 accumulator ready", 9: "tile0 epilogue done", 10: "tile1 accumulator ready", 11: "tile1 epilogue done", 15: "exit"}
if std:
    names.update({3: "epilogue chunk 0: TMEM loaded", 4: "epilogue chunk 0: staged in smem", 5: "epilogue chunk 0: stored",
                  6: "epilogue chunk 1: stored"})
print("traced GEMM, last launch of a B=%d beam-3 call with T=%d: us since the first CTA's entry" % (B, steps))
for k in sorted(names):
    col = t[:, k]
    ok = col > 0
    if ok.any():
        v = (col[ok] - t0) / 1e3
        print("  %-28s n=%3d  median %6.1f  min %6.1f  max %6.1f" % (names[k], ok.sum(), np.median(v), v.min(), v.max()))
