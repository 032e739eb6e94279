"""Times the training iterations of BASELINE configs[2] / [3] on synthetic data (one process per GPU; torchrun for N > 1):
  xe : train_xe.py:160-192  — xe + domain-alignment + seq2seq losses, backward, gradient all-reduce, clamp + Adam
  rl : models/decoder.py:62-170 — sampled (x samples_per_image) + greedy decode, device CIDEr-D reward, REINFORCE + DA + XE
Usage: python profiles/train_bench.py xe|rl [batch] [iters] [samples_per_image]
Prints one JSON line on rank 0 (ms per iteration = max over ranks of CUDA-event time; rows/s over all ranks)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from insenticap_model_b200 import reward as R  # noqa: E402
from insenticap_model_b200 import synthetic as syn  # noqa: E402
from insenticap_model_b200 import train as TR  # noqa: E402
from insenticap_model_b200.captioner import Captioner  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "xe"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
spi = int(sys.argv[4]) if len(sys.argv) > 4 else 5
V, T = 10000, 16
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

m = Captioner(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS), precision="bf16x3")
m.load_state_dict(syn.synthetic_state_dict(V, 0))
m = m.to(dev).train()
optim = TR.FusedClampAdam(m, lr=4e-4, grad_clip=0.1)
g = torch.Generator(device=dev).manual_seed(100 + rank)  # every rank holds its own images
fc = torch.rand(B, 2048, device=dev, generator=g)
att = torch.rand(B, 14, 14, 2048, device=dev, generator=g)
cpts = torch.randint(4, V, (B, 5), device=dev, generator=g)
sentis = torch.randint(4, V, (B, 10), device=dev, generator=g)
labels = (torch.arange(B, device=dev) % 3).long()
caps = torch.randint(4, V, (B, T + 1), device=dev, generator=g)
caps[:, 0] = 1
lengths = [T] * B
s2s = (caps, lengths, cpts, sentis, labels)
if what == "rl":
    refs = syn.synthetic_references(B, V, 5, seed=3 + rank)
    fns = ["img%d" % i for i in range(B)]
    gts = {fn: refs[i] for i, fn in enumerate(fns)}
    scorer = R.get_ciderd_scorer({"train": gts}, 1, 2, device=dev)
    scorer.register_ground_truth(fns, gts, 1, 2)
    batch = (fns, fc, att, caps, lengths, cpts, sentis, labels, gts)
    step = lambda: TR.rl_iteration(m, optim, scorer, batch, max_seq_len=T, seq2seq_batch=s2s, samples_per_image=spi)
    rows = B * (spi + 1)
else:
    batch = (fc, att, caps, lengths, cpts, labels)
    step = lambda: TR.xe_iteration(m, optim, batch, s2s)
    rows = B

def timed(n):
    for _ in range(2):
        o = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    e0.record()
    h0 = time.perf_counter()
    for _ in range(n):
        o = step()
    host_issue_ms[0] = (time.perf_counter() - h0) * 1e3 / n  # host time to ISSUE an iteration (no sync inside the loop)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), o


host_issue_ms = [None]


e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ms_val, out = timed(iters)  # gradient all-reduce in buckets started inside the backward (the default)
ms = torch.tensor([ms_val], device=dev, dtype=torch.float64)
note = optim.overlap_note
comm = {"allreduce_ms": None, "allreduce_busbw_gbs": None, "ms_per_iteration_no_overlap": None, "overlap": optim.overlap_note}
if world > 1:
    optim.overlap = False  # one blocking all-reduce of the flat gradient after the backward (round 1's scheme)
    comm["ms_per_iteration_no_overlap"], _ = timed(iters)
    optim.overlap = True
    ms_val, out = timed(iters)  # again, after the comparison run: order effects (allocator, clocks) show up as a difference
    comm["ms_per_iteration_overlap_rerun"] = ms_val
    comm["overlap"] = note
    buf = torch.zeros_like(optim.flat_g)
    for _ in range(2):
        dist.all_reduce(buf)
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(5):
        dist.all_reduce(buf)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 5], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    comm["allreduce_ms"] = float(t.item())
    comm["allreduce_busbw_gbs"] = 2.0 * (world - 1) / world * buf.numel() * 4 / (float(t.item()) * 1e-3) / 1e9
    comm["grad_bytes"] = buf.numel() * 4
    del buf
chk = optim.flat_p.double().sum().reshape(1)
if world > 1:
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    in_sync = bool((hi - lo).abs().item() == 0.0)
else:
    in_sync = True
if rank == 0:
    print(json.dumps({"workload": what, "n_gpus": world, "batch_per_gpu": B, "samples_per_image": spi if what == "rl" else None,
                      "ms_per_iteration": float(ms.item()), "host_issue_ms_per_iteration": host_issue_ms[0], "rows_per_s": world * rows / (float(ms.item()) * 1e-3),
                      "losses": {k: float(v) for k, v in out.items()}, "replicas_in_sync": in_sync,
                      "peak_mem_GB": torch.cuda.max_memory_allocated(dev) / 1e9, **comm}), flush=True)
# per-kernel-class device time of one more iteration (CUDA events around every launch of the library)
if rank == 0 and os.environ.get("ISC_TRAIN_PROFILE"):
    import ctypes as C
    from insenticap_model_b200 import _lib
    lib = _lib.load()
    lib.isc_profile_reset()
    lib.isc_profile_enable(1)
    n0 = lib.isc_launch_count()
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    lib.isc_profile_enable(0)
    out = {}
    for i, name in enumerate(_lib.KERNEL_CLASSES):
        tm, wk, n = C.c_double(), C.c_double(), C.c_int64()
        lib.isc_profile_read(i, C.byref(tm), C.byref(wk), C.byref(n))
        if n.value:
            out[name] = {"ms": round(tm.value, 3), "launches": n.value}
    print(json.dumps({"profiled_iteration_ms": e0.elapsed_time(e1), "library_launches": lib.isc_launch_count() - n0, "classes": out}))
if world > 1:
    dist.destroy_process_group()
