"""Multi-GPU plumbing of the rollout path: contiguous sharding of start states and the two
Allreduce steps of `mpi_statistics_scalar` (utilities/mpi_tools.py:71-92).

Paths are independent units, so there is no data-path collective: every rank rolls out its own
block of start states (Philox counters use GLOBAL path ids, so results do not depend on the
sharding), and only the five running sums (n, sum adv, sum cadv, sum ret, sum cret) and then the
sum of squared deviations cross ranks.  One process per GPU, `torch.distributed` (NCCL on GPUs,
gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous block partition [lo, hi) of n items; sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def make_reduce_fn(group=None):
    """In-place SUM all-reduce of a small tensor, or None for a single process."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None

    def reduce_fn(t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t
    return reduce_fn


def combine_pass1(sums):
    """sums = (n, sum_adv, sum_cadv, sum_ret, sum_cret) after the first all-reduce ->
    float32 means exactly as mpi_statistics_scalar forms them (sum and n cast to float32 first)."""
    n = np.float32(sums[0])
    if sums[0] == 0:
        z = np.float32(0)
        return dict(n=0, adv_mean=z, cadv_mean=z, ret_mean=z, cret_mean=z)
    return dict(n=int(sums[0]), adv_mean=np.float32(sums[1]) / n, cadv_mean=np.float32(sums[2]) / n,
                ret_mean=np.float32(np.float32(sums[3]) / n), cret_mean=np.float32(np.float32(sums[4]) / n))


def combine_pass2(ssq, n):
    """Population std from the all-reduced sum of squared deviations (mpi_tools.py:85-86)."""
    return np.sqrt(np.float32(ssq) / np.float32(n))
