"""Batch sharding across the GPUs of one box.

Every IK problem is independent, so the multi-GPU path has NO data-path collective: rank r of G solves the
contiguous slice [r*B/G, (r+1)*B/G) of the batch with the same kernel and the same constants.  The only
communication is the optional gather of the results to rank 0 (torch.distributed all_gather; NCCL over NVLink on
the GPU box, gloo in the CPU tests) and the timing reduction in bench.py.
"""
import numpy as np


def shard_range(B, rank, world):
    """Contiguous, balanced slice of a batch of B problems for `rank` of `world`."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    lo = (B * rank) // world
    hi = (B * (rank + 1)) // world
    return lo, hi


def shard_sizes(B, world):
    return [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]


def gather_results(local, B, dist=None):
    """Gather per-rank result arrays (dict name -> tensor with the batch as LAST dim) to every rank.

    `local` holds this rank's slice; slices may differ in size by one, so they are padded to the largest shard
    for the all_gather and trimmed afterwards.  With dist=None (single process) the input is returned as is."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch

    world = dist.get_world_size()
    sizes = shard_sizes(B, world)
    pad_to = max(sizes)
    out = {}
    for name, t in local.items():
        pad = pad_to - t.shape[-1]
        tp = torch.nn.functional.pad(t, (0, pad)) if pad else t
        tp = tp.contiguous()
        bufs = [torch.empty_like(tp) for _ in range(world)]
        dist.all_gather(bufs, tp)
        out[name] = torch.cat([b[..., :n] for b, n in zip(bufs, sizes)], dim=-1)
    return out


def reduce_throughput(units, seconds, dist=None):
    """Whole-job throughput: units of all ranks / max-over-ranks time."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return units / seconds, seconds
    import torch

    t = torch.tensor([seconds], dtype=torch.float64)
    u = torch.tensor([float(units)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return u.item() / t.item(), t.item()
