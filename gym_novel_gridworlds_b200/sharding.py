"""Multi-GPU plumbing: environments are independent, so a job of N envs is cut into contiguous global-id ranges,
one per rank (one process per GPU).  No collective is on the step path; the only exchange is the all-reduce of the
8-double episode-statistics vector (NCCL over NVLink on GPUs, gloo in the CPU tests) and the max-over-ranks of timings."""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """Contiguous [begin, end) of global env ids owned by `rank`; sizes differ by at most one 32-env tile."""
    tiles = (n_total + 31) // 32
    per, extra = divmod(tiles, world)
    t0 = rank * per + min(rank, extra)
    t1 = t0 + per + (1 if rank < extra else 0)
    return min(t0 * 32, n_total), min(t1 * 32, n_total)


def allreduce_stats(stats):
    """Sum the per-rank statistics vector (float64[8], order = opcodes.STAT_NAMES) over all ranks, in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def max_over_ranks(value, device=None):
    """Device-timed durations are reported as the max over ranks."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
