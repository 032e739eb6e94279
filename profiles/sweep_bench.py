"""BASELINE configs[4]: throughput sweep of Captioner.beam_search over beam sizes 1 / 3 / 5 and batch sizes 64 .. 8192 per
GPU (V = 10000, 16 tokens, CUDA-graph replay, device-resident inputs), on 1 .. 8 B200 (one process per GPU under torchrun,
every rank its own images: weak scaling, no data-path collective; device time, max over ranks), with the reference's CPU
path (per-image Captioner.sample on the host cores, bench.CpuReference) beside it.
Usage: [torchrun --nproc-per-node N] python profiles/sweep_bench.py [precision] [out.json]
Prints one JSON line per (beam, batch), a markdown table, and writes the rows to out.json on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from insenticap_model_b200 import synthetic as syn  # noqa: E402
from insenticap_model_b200.captioner import Captioner  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
out_path = sys.argv[2] if len(sys.argv) > 2 else None
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
V, T = 10000, 16
BATCHES, BEAMS = (64, 256, 1024, 4096, 8192), (1, 3, 5)
m = Captioner(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS), precision=precision)
m.load_state_dict(syn.synthetic_state_dict(V, 0))
m = m.to(dev).eval()
m.use_cuda_graph = True
m.graph_outputs_fresh = False


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


rows = []
for B in BATCHES:
    g = torch.Generator(device=dev).manual_seed(B + 17 * rank)
    fc = torch.rand(B, 2048, device=dev, generator=g)
    att = torch.rand(B, 14, 14, 2048, device=dev, generator=g)
    sentis = torch.randint(4, V, (B, 10), device=dev, generator=g)
    labels = (torch.arange(B, device=dev) % 3).long()
    for K in BEAMS:
        m._graphs.clear()
        step = lambda: m.beam_search(fc, att, sentis, labels, beam_size=K, max_seq_len=T)
        for _ in range(3):
            step()
        n = max(3, min(20, int(2e4 / (B * K) ** 0.9) + 3))
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        r = {"beam": K, "batch_per_gpu": B, "n_gpus": world, "ms_per_call": ms, "captions_per_s": world * B / (ms * 1e-3),
             "precision": precision}
        rows.append(r)
        if rank == 0:
            print(json.dumps(r), flush=True)
    del fc, att
    m._graphs.clear()
    m._ws.clear()
    torch.cuda.empty_cache()

cpu = {}
if rank == 0:
    # the reference's own path: one image at a time (its only beam search), all host cores; cost per image is independent
    # of the batch size, so one bounded sample per beam size
    from bench import CpuReference  # noqa: E402
    for K in BEAMS:
        ref = CpuReference(8, K, os.cpu_count())
        ref.rate(1)
        rate, dt, threads = ref.rate(6)
        cpu[K] = {"captions_per_s": rate, "cores": threads, "kind": ref.kind, "sample": "6 images, %.1f s" % dt}
        print(json.dumps({"cpu_reference_beam": K, **cpu[K]}), flush=True)
    print("\n| batch/GPU | beam 1 | beam 3 | beam 5 |  (captions/s, %s, %d GPU)\n|---|---|---|---|" % (precision, world))
    for B in BATCHES:
        print("| %d | " % B + " | ".join("%.0f" % next(r["captions_per_s"] for r in rows if r["batch_per_gpu"] == B and r["beam"] == K)
                                          for K in BEAMS) + " |")
    print("| CPU reference (%d cores) | " % cpu[1]["cores"] + " | ".join("%.1f" % cpu[K]["captions_per_s"] for K in BEAMS) + " |")
    if out_path:
        json.dump({"rows": rows, "cpu_reference": {str(k): v for k, v in cpu.items()}}, open(out_path, "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
