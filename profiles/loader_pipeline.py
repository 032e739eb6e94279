"""Stage timeline of shard -> DevicePrefetcher -> Captioner.beam_search on a fp16 feature shard:
per batch, when the worker thread's gather started / ended, when its H2D copies were issued, when the copy landed (event),
when the consumer got the batch and when its decode finished. Usage: python profiles/loader_pipeline.py [N] [B] [depth]"""
import os
import sys
import tempfile
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from insenticap_model_b200 import dataloader as dl  # noqa: E402
from insenticap_model_b200 import synthetic as syn
from insenticap_model_b200.captioner import Captioner

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 3
pin = len(sys.argv) > 4 and sys.argv[4] == "pin"  # page-lock the shard: batches go records -> HBM without the staging gather
V = 10000
root = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
d = tempfile.mkdtemp(dir=root)
names = ["img%06d" % i for i in range(N)]
g = torch.Generator().manual_seed(0)
fc = torch.rand(N, 2048, generator=g)
att = torch.rand(N, 196, 2048, generator=g)
_, _, cpts, sentis, labels = syn.synthetic_inputs(N, V, seed=1)
path = dl.FeatureShard.write(os.path.join(d, "f16.iscf"), names, fc, att, dtype="fp16")
del fc, att
log = []
T0 = [0.0]


def stamp(what, **kw):
    log.append((time.perf_counter() - T0[0], threading.current_thread().name, what, kw))


try:
    sh = dl.FeatureShard(path)
    orig_gather = sh.gather

    def gather(indices, **kw):
        stamp("gather>", n=len(indices), att=kw.get("want_att", True))
        r = orig_gather(indices, **kw)
        stamp("gather<")
        return r

    sh.gather = gather
    orig_c2d = sh.copy_to_device

    def c2d(indices, **kw):
        stamp("copy_to_device>", n=len(indices), att=kw.get("want_att", True))
        r = orig_c2d(indices, **kw)
        stamp("copy_to_device issued")
        return r

    sh.copy_to_device = c2d
    if pin:
        print("page-locked:", sh.pin("cuda:0"))
    orig_to = dl._to_device

    def to_dev(x, dev):
        top = isinstance(x, tuple) and len(x) == 6
        if top:
            stamp("h2d>")
        r = orig_to(x, dev)
        if top:
            stamp("h2d issued")
        return r

    dl._to_device = to_dev
    concepts = {fn: cpts[i].tolist() for i, fn in enumerate(names)}
    sentiments = {fn: sentis[i].tolist() for i, fn in enumerate(names)}
    labs = [(fn, int(labels[i])) for i, fn in enumerate(names)]
    m = Captioner(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS), precision="bf16x3")
    m.load_state_dict(syn.synthetic_state_dict(V, 0))
    m = m.cuda().eval()
    loader = dl.get_rl_senti_dataloader(sh, sh, concepts, sentiments, labs, 0, 5, 10, batch_size=B, shuffle=True)
    with torch.no_grad():
        runs = []
        for epoch in range(6):
            torch.cuda.synchronize()
            log.clear()
            T0[0] = time.perf_counter()
            outs = []
            for fns, f, a, c, s, l in dl.DevicePrefetcher(loader, "cuda:0", depth=depth):
                stamp("consumer got batch")
                outs.append(m.beam_search(f, a, s, l, 3, 1, 16)[0])
                stamp("decode issued")
            torch.cuda.synchronize()
            dt = time.perf_counter() - T0[0]
            print("epoch %d: %d images in %.1f ms = %.0f captions/s" % (epoch, N, dt * 1e3, N / dt))
            if epoch > 0:
                runs.append((dt, epoch, list(log)))
        dt, epoch, slow = max(runs)
        print("timeline of the slowest epoch after the first (%d):" % epoch)
        for t, th, what, kw in slow:
            print("%8.2f ms  %-12s %s %s" % (t * 1e3, th[:12], what, kw or ""))
finally:
    os.remove(path)
    os.rmdir(d)
