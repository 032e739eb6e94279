"""Times SentimentDetector.sample (two 3x3 convolutions as tcgen05 GEMMs + head) on one B200: images/s and the
convolution GEMMs' share. Usage: python profiles/senti_bench.py [batch]"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from insenticap_model_b200 import _lib  # noqa: E402
from insenticap_model_b200 import synthetic as syn  # noqa: E402
from insenticap_model_b200.sentiment_detector import SentimentDetector  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = SentimentDetector(syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS, sentiment_convs_num=2, sentiment_fcs_num=2))
m.load_state_dict(syn.senti_detector_state_dict(0))
m = m.cuda().eval()
att = torch.rand(B, 14, 14, 2048, device="cuda")
for _ in range(2):
    m.sample(att, 0.7)
torch.cuda.synchronize()
lib = _lib.load()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    m.sample(att, 0.7)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
lib.isc_profile_reset()
lib.isc_profile_enable(1)
m.sample(att, 0.7)
torch.cuda.synchronize()
lib.isc_profile_enable(0)
tm, wk, cnt = C.c_double(), C.c_double(), C.c_int64()
lib.isc_profile_read(0, C.byref(tm), C.byref(wk), C.byref(cnt))
print(json.dumps({"batch": B, "ms_per_call": ms, "images_per_s": B / (ms * 1e-3), "conv_gemm_ms": tm.value,
                  "conv_gemm_tflops_x3": wk.value / (tm.value * 1e-3) / 1e12, "gflop_per_image_x1": wk.value / 3 / B / 1e9}))
