"""Multi-GPU partitioning of the NTT workload (host logic only).

The path has no exchange step: polynomials (batch index) and RNS limbs are
independent.  One process per GPU; each rank transforms its shard
on its own context and nothing crosses NVLink on the data path.  The helpers
here decide who owns what and reduce the timing (max over ranks); they work
with any torch.distributed backend (NCCL on the GPU box, gloo in the CPU
tests).
"""


def split_range(total, world, rank):
    """Contiguous, balanced [begin, end) share of `total` units for `rank`."""
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def limb_shard(limbs, world, rank):
    """Limb range of `rank` when an RNS basis is sharded by prime (BASELINE
    configs[2]: 32 limbs -> 32/16/8/4 per GPU).  Contiguous so that a rank
    uploads the twiddle tables of its own primes only."""
    return split_range(limbs, world, rank)


def batch_shard(batch, world, rank):
    """Batch range of `rank` when polynomials are sharded by batch index
    (BASELINE configs[3] and [4])."""
    return split_range(batch, world, rank)


def gather_limb_sharded(local, limbs, n, batch, dist):
    """Optional result gather (NOT on the hot path): every rank holds
    [batch][own limbs][n]; returns the full [batch][limbs][n] array on every
    rank.  `dist` is torch.distributed (initialised)."""
    import numpy as np
    import torch
    world = dist.get_world_size()
    pieces = [None] * world
    dist.all_gather_object(pieces, np.ascontiguousarray(local))
    out = np.empty(batch * limbs * n, dtype=local.dtype).reshape(batch, limbs, n)
    for r, piece in enumerate(pieces):
        lo, hi = limb_shard(limbs, world, r)
        out[:, lo:hi, :] = piece.reshape(batch, hi - lo, n)
    return out.reshape(-1)


def max_over_ranks(value, dist, device=None):
    """Timing reduction of the bench contract: the slowest rank decides."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
