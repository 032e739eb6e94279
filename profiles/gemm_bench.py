"""Times the tcgen05 GEMM on the decode path's shapes through isc_gemm_tn (CUDA events around the GEMM launch only,
via the library's profile hooks). Usage: python profiles/gemm_bench.py [bf16x3|bf16]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from insenticap_model_b200 import _lib  # noqa: E402

prec_name = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
prec = _lib.PRECISIONS[prec_name]
passes = 3 if prec_name == "bf16x3" else 1
lib = _lib.load()
SHAPES = [("gates (att/lang LSTM)", 3072, 2048, 1536), ("h projections", 3072, 1536, 512), ("gate", 3072, 512, 1024),
          ("logits", 3072, 10000, 512), ("att_embed chunk 96", 18816, 512, 2048), ("att2att chunk 96", 18816, 512, 512),
          ("att_embed chunk 64", 12544, 512, 2048), ("square 8192", 8192, 8192, 8192)]
for name, M, N, K in SHAPES:
    A = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    out = torch.empty(M, N, device="cuda")
    ws = torch.empty(lib.isc_gemm_workspace_bytes(prec, M, N, K), dtype=torch.uint8, device="cuda")

    def run():
        _lib.check(lib.isc_gemm_tn(prec, _lib.ptr(A), K, _lib.ptr(W), K, None, _lib.ptr(out), N, M, N, K, 0, _lib.ptr(ws),
                                   ws.numel(), _lib.stream_ptr()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    lib.isc_profile_reset()
    lib.isc_profile_enable(1)
    n = 10
    for _ in range(n):
        run()
    torch.cuda.synchronize()
    lib.isc_profile_enable(0)
    tm, wk, cnt = C.c_double(), C.c_double(), C.c_int64()
    _lib.check(lib.isc_profile_read(0, C.byref(tm), C.byref(wk), C.byref(cnt)))
    lib.isc_profile_reset()
    us = 1e3 * tm.value / max(cnt.value, 1)
    print("%-24s M=%6d N=%6d K=%5d  %8.1f us  %7.1f TFLOP/s (x%d passes: %7.1f)" % (
        name, M, N, K, us, 2.0 * M * N * K / us / 1e6, passes, 2.0 * M * N * K * passes / us / 1e6), flush=True)
