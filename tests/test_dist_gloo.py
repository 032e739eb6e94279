"""Host-side multi-process logic (world_size 2, gloo, CPU): contiguous image sharding with no data-path
collective and the final gather. The decode itself is stubbed by a per-row function — the CUDA path has
no CPU implementation by design."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from insenticap_model_b200 import dist as D


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    fc = torch.rand(n, 6, generator=g)
    att = torch.rand(n, 3, 6, generator=g)
    labels = torch.arange(n) % 3
    calls = []

    def fake_decode(fc_l, att_l, sw_l, sl_l):
        calls.append(fc_l.shape[0])
        assert sw_l is None
        tok = (fc_l.sum(1, keepdim=True) * 100).long().repeat(1, 4) + sl_l.view(-1, 1)
        score = att_l.double().sum((1, 2)).view(-1, 1)
        return tok, score

    tok, score = D.sharded_decode(fake_decode, [fc, att, None, labels])
    a, b = D.shard_range(n, rank, world)
    ok = calls == [b - a]
    ok &= torch.equal(tok, (fc.sum(1, keepdim=True) * 100).long().repeat(1, 4) + labels.view(-1, 1))
    ok &= torch.equal(score, att.double().sum((1, 2)).view(-1, 1))
    out_q.put((rank, bool(ok), tuple(tok.shape)))
    dist.destroy_process_group()


def test_sharded_decode_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 9  # uneven split: 5 + 4 images
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, (n, 4)), (1, True, (n, 4))]


# ---------------------------------------------------------------------------------------------------------
# Training exchange step: the flat-gradient all-reduce of insenticap_model_b200.train (world 2, gloo). The
# clamp + Adam kernel that follows is CUDA-only; here a tiny linear model checks that the averaged flat gradient
# equals the gradient of the single large batch (equal per-rank batch sizes, mean losses).
def _train_worker(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from insenticap_model_b200 import train as TR
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 3))
    ref = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 3))
    ref.load_state_dict(net.state_dict())
    flat_p, flat_g = TR.flatten_parameters(net)
    ok = all(p.data_ptr() >= flat_p.data_ptr() for p in net.parameters())  # parameters are views of the flat buffer
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 5, generator=g), torch.randn(8, 3, generator=g)
    a, b = D.shard_range(8, rank, world)
    torch.nn.functional.mse_loss(net(x[a:b]), y[a:b]).backward()
    ok &= all(p.grad.data_ptr() >= flat_g.data_ptr() and p.grad.data_ptr() < flat_g.data_ptr() + flat_g.numel() * 4
              for p in net.parameters())  # autograd accumulated in place into the flat gradient views
    w = TR.allreduce_gradients(flat_g)
    torch.nn.functional.mse_loss(ref(x), y).backward()
    want = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
    ok &= w == world and torch.allclose(flat_g / w, want, atol=1e-6)
    out_q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
