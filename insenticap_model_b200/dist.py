"""Multi-GPU plumbing for the decode path: images are independent units, so a global batch is split
contiguously by image across ranks (beams of one image never straddle GPUs), every rank decodes its
shard with NO data-path collective, and results are gathered once at the end.
One process per GPU, torch.distributed (NCCL on GPUs; gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous [start, end) of `n` items owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank: int, world: int):
    """Slice every tensor (or None) of a batch along dim 0 to this rank's image range."""
    n = next(t.shape[0] for t in tensors if t is not None)
    a, b = shard_range(n, rank, world)
    return [None if t is None else t[a:b] for t in tensors]


def all_gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Concatenate per-rank row blocks (uneven sizes allowed) back into the global [n_total, ...]."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    max_rows = max(b - a for a, b in sizes)
    pad = local.new_zeros((max_rows,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: b - a] for r, (a, b) in enumerate(sizes)], dim=0)


def sharded_decode(decode_fn, batch, group=None):
    """Run `decode_fn(*local_shard) -> tuple of tensors [n_local, ...]` on this rank's images and gather
    every output to all ranks. `batch` is the GLOBAL batch (each rank holds or can index it)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = next(t.shape[0] for t in batch if t is not None)
    outs = decode_fn(*shard_batch(batch, rank, world))
    if world == 1:
        return outs
    return tuple(all_gather_rows(o.contiguous(), n, group) for o in outs)


def beam_search_sharded(captioner, fc_feats, att_feats, senti_words=None, senti_labels=None, beam_size=3,
                        decoding_constraint=1, max_seq_len=16, group=None):
    """Captioner.beam_search over a global batch split across the ranks of `group`."""
    dev = captioner.classifier.weight.device

    def fn(fc, att, sw, sl):
        mv = lambda t: None if t is None else t.to(dev, non_blocking=True)
        return captioner.beam_search(mv(fc), mv(att), mv(sw), mv(sl), beam_size, decoding_constraint, max_seq_len)

    return sharded_decode(fn, [fc_feats, att_feats, senti_words, senti_labels], group)
