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


def e2e_plan(rates, limbs, batch):
    """Who moves what in the end-to-end (host -> GPU -> host) path.

    `rates[r]` is rank r's measured host <-> device rate while all ranks copy.
    With equal limb shards the job ends with the slowest rank, and on a box
    whose GPUs get unequal shares of the host fabric that wastes the others'
    bandwidth.  So ranks are paired, slowest with fastest: a pair shares the
    limbs of both equal shards and splits the batch entries in proportion to
    the two rates (at least one entry each; equal rates give equal halves, i.e.
    as many polynomials per rank as the equal limb shards).

    Returns, for every rank, (limb list, batch begin, batch end).  One rank, or
    an odd number of ranks: the plain limb shards over the whole batch."""
    world = len(rates)
    if world == 1 or world % 2 or batch < 2:
        return [(list(range(*limb_shard(limbs, world, r))), 0, batch)
                for r in range(world)]
    order = sorted(range(world), key=lambda r: (rates[r], r))
    plan = [None] * world
    for i in range(world // 2):
        slow, fast = order[i], order[world - 1 - i]
        both = sorted(list(range(*limb_shard(limbs, world, slow)))
                      + list(range(*limb_shard(limbs, world, fast))))
        total = rates[slow] + rates[fast]
        share = rates[slow] / total if total > 0 else 0.5
        cut = min(batch - 1, max(1, int(round(batch * share))))
        plan[slow] = (both, 0, cut)
        plan[fast] = (both, cut, batch)
    return plan


def chunk_sizes(count, chunks):
    """`count` batch entries cut into at most `chunks` slices of whole entries,
    sizes differing by at most one, larger ones first"""
    chunks = max(1, min(chunks, count))
    base, extra = divmod(count, chunks)
    return [base + (1 if c < extra else 0) for c in range(chunks)]


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
