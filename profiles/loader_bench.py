"""Input pipeline throughput on the GPU box: native shard gather into pinned memory
(GB/s by thread count), and images/s of shard -> DevicePrefetcher -> Captioner.beam_search (beam 3, 16 tokens, V=10000)
for an fp32 and a fp16 shard in the token-exact mode and a bf16 shard in the bf16 mode. Usage: python profiles/loader_bench.py [N] [B]"""
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from insenticap_model_b200 import dataloader as dl  # noqa: E402
from insenticap_model_b200 import synthetic as syn
from insenticap_model_b200.captioner import Captioner

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
V = 10000
root = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
d = tempfile.mkdtemp(dir=root)
names = ["img%06d" % i for i in range(N)]
g = torch.Generator().manual_seed(0)
fc = torch.rand(N, 2048, generator=g)
att = torch.rand(N, 196, 2048, generator=g)
_, _, cpts, sentis, labels = syn.synthetic_inputs(N, V, seed=1)
paths = {k: dl.FeatureShard.write(os.path.join(d, k + ".iscf"), names, fc, att, dtype=k) for k in ("fp32", "fp16", "bf16")}
del fc, att
try:
    for k in ("fp32", "fp16", "bf16"):
        sh = dl.FeatureShard(paths[k])
        idx = torch.randperm(N, generator=g)[:B].tolist()
        gb = B * (1 + 196) * 2048 * (4 if k == "fp32" else 2) / 1e9
        for threads in (1, 4, 8, 16, 32):
            sh.gather(idx, threads=threads)
            t0 = time.perf_counter()
            for _ in range(5):
                sh.gather(idx, threads=threads)
            dt = (time.perf_counter() - t0) / 5
            print("gather %s B=%d threads=%2d: %.2f ms, %.1f GB/s, %.0f images/s" % (k, B, threads, dt * 1e3, gb / dt, B / dt))
    concepts = {fn: cpts[i].tolist() for i, fn in enumerate(names)}
    sentiments = {fn: sentis[i].tolist() for i, fn in enumerate(names)}
    labs = [(fn, int(labels[i])) for i, fn in enumerate(names)]
    for k, prec, pin in (("fp32", "bf16x3", False), ("fp16", "bf16x3", False), ("bf16", "bf16", False), ("fp16", "bf16x3", True)):
        m = Captioner(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS), precision=prec)
        m.load_state_dict(syn.synthetic_state_dict(V, 0))
        m = m.cuda().eval()
        sh = dl.FeatureShard(paths[k])
        if pin and not sh.pin("cuda:0"):
            print("cannot page-lock the mapping on this platform")
            continue
        loader = dl.get_rl_senti_dataloader(sh, sh, concepts, sentiments, labs, 0, 5, 10, batch_size=B, shuffle=True)
        with torch.no_grad():
            for epoch in range(6):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                outs = []
                for fns, f, a, c, s, l in dl.DevicePrefetcher(loader, "cuda:0", depth=3):
                    outs.append(m.beam_search(f, a, s, l, 3, 1, 16)[0])
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                print("shard(%s%s) -> prefetch -> beam-3 decode (%s), epoch %d: %d images in %.1f ms = %.0f captions/s"
                      % (k, ", page-locked" if pin else "", prec, epoch, N, dt * 1e3, N / dt))
finally:
    for p in paths.values():
        os.remove(p)
    os.rmdir(d)
