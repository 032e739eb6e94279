"""Bandwidth and jitter of isc_shard_copy_to_device alone (page-locked fp16 shard, 512 shuffled records per call):
one cudaMemcpyAsync per record (default) against the zero-copy gather kernel (ISC_SHARD_COPY=kernel).
Usage: python profiles/shard_copy_bench.py [N] [B]"""
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from insenticap_model_b200 import dataloader as dl  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
root = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
d = tempfile.mkdtemp(dir=root)
g = torch.Generator().manual_seed(0)
path = dl.FeatureShard.write(os.path.join(d, "f16.iscf"), ["i%d" % i for i in range(N)], torch.rand(N, 2048, generator=g),
                             torch.rand(N, 196, 2048, generator=g), dtype="fp16")
try:
    sh = dl.FeatureShard(path)
    assert sh.pin("cuda:0")
    gb = B * 197 * 2048 * 2 / 1e9
    times = []
    for it in range(30):
        idx = torch.randperm(N, generator=g)[:B].tolist()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fc, att = sh.copy_to_device(idx)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        times.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
    tot = sorted(t[1] for t in times[3:])
    print("copy_to_device of %d fp16 records (%.2f GB): issue %.2f ms median, done %.2f ms median (%.1f GB/s), min %.2f, max %.2f ms"
          % (B, gb, sorted(t[0] for t in times[3:])[len(tot) // 2], tot[len(tot) // 2], gb / (tot[len(tot) // 2] * 1e-3), tot[0], tot[-1]))
    print("all (ms):", " ".join("%.1f" % t[1] for t in times))
finally:
    os.remove(path)
    os.rmdir(d)
