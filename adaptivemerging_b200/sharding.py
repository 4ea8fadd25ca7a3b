"""Multi-GPU plumbing for batched independent scenes (config B of BASELINE.json).

Independent scene instances shard across GPUs with NO data-path collective: rank r owns a contiguous block of
scene ids and steps it in its own am3d context.  torch.distributed is only used for the barrier that brackets the
timed region and for reducing the timing (MAX) and the processed-unit counters (SUM).
"""


def shard_scenes(total_scenes, rank, world):
    """Contiguous block [start, start + count) of scene ids owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(total_scenes, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def reduce_stats(maxes, sums, dist=None, device="cpu"):
    """(element-wise MAX over ranks of `maxes` - timed spans -, element-wise SUM over ranks of `sums` - processed units)."""
    maxes, sums = [float(x) for x in maxes], [float(x) for x in sums]
    if dist is None:
        return maxes, sums
    import torch
    t = torch.tensor(maxes, dtype=torch.float64, device=device)
    u = torch.tensor(sums, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return t.tolist(), u.tolist()


def reduce_step_stats(ms, units, dist=None, device="cpu"):
    """(max over ranks of the timed span in ms, sum over ranks of the units processed)."""
    t, u = reduce_stats([ms], [units], dist, device)
    return t[0], u[0]
