"""Where the HOST time of an XE training iteration goes (issue time, no sync inside): time spent inside the library's C calls
(kernel launches, tensor-map encodes) against the Python / torch time around them.
Usage: python profiles/train_host_profile.py [batch] [iters]"""
import collections
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from insenticap_model_b200 import _lib  # noqa: E402
from insenticap_model_b200 import synthetic as syn  # noqa: E402
from insenticap_model_b200 import train as TR  # noqa: E402
from insenticap_model_b200.captioner import Captioner  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
V, T = 10000, 16
dev = torch.device("cuda", 0)
m = Captioner(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS), precision="bf16x3")
m.load_state_dict(syn.synthetic_state_dict(V, 0))
m = m.to(dev).train()
optim = TR.FusedClampAdam(m, lr=4e-4, grad_clip=0.1)
g = torch.Generator(device=dev).manual_seed(100)
fc = torch.rand(B, 2048, device=dev, generator=g)
att = torch.rand(B, 14, 14, 2048, device=dev, generator=g)
cpts = torch.randint(4, V, (B, 5), device=dev, generator=g)
sentis = torch.randint(4, V, (B, 10), device=dev, generator=g)
labels = (torch.arange(B, device=dev) % 3).long()
caps = torch.randint(4, V, (B, T + 1), device=dev, generator=g)
caps[:, 0] = 1
lengths = [T] * B
s2s = (caps, lengths, cpts, sentis, labels)
batch = (fc, att, caps, lengths, cpts, labels)
step = lambda: TR.xe_iteration(m, optim, batch, s2s)

lib = _lib.load()
spent = collections.defaultdict(lambda: [0, 0.0])


class Timed:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, *a):
        t = time.perf_counter()
        r = self.fn(*a)
        e = spent[self.name]
        e[0] += 1
        e[1] += time.perf_counter() - t
        return r


class Proxy:
    def __init__(self, lib):
        object.__setattr__(self, "_l", lib)
        object.__setattr__(self, "_c", {})

    def __getattr__(self, k):
        c = object.__getattribute__(self, "_c")
        if k not in c:
            f = getattr(object.__getattribute__(self, "_l"), k)
            c[k] = Timed(k, f) if callable(f) else f
        return c[k]


for _ in range(3):
    step()
torch.cuda.synchronize()
_lib._lib = Proxy(lib)  # load() now hands out the timing proxy
# pure host cost: ONE iteration issued into an empty launch queue (a queue that the GPU drains more slowly than the host
# fills it blocks the host inside cudaLaunchKernel, and the host time then just mirrors the device time)
free = []
for _ in range(5):
    torch.cuda.synchronize()
    spent.clear()
    t0 = time.perf_counter()
    step()
    free.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
print("host ms for one iteration issued into an empty queue: " + ", ".join("%.2f" % x for x in free))
for k, (n, s_) in sorted(spent.items(), key=lambda kv: -kv[1][1])[:6]:
    print("  C call %-32s n %4d  ms %7.3f" % (k, n, s_ * 1e3))
spent.clear()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
for _ in range(iters):
    step()
pr.disable()
host = (time.perf_counter() - t0) / iters * 1e3
torch.cuda.synchronize()
print("host issue ms / iteration (under cProfile): %.2f" % host)
for k, (n, s) in sorted(spent.items(), key=lambda kv: -kv[1][1])[:12]:
    print("  C call %-32s n/iter %6.1f  ms/iter %7.3f" % (k, n / iters, s / iters * 1e3))
st = pstats.Stats(pr)
st.sort_stats("tottime")
st.print_stats(18)
