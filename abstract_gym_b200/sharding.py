"""Multi-GPU layout of the path: environments shard as contiguous slices, one process per GPU,
replicated (or co-sharded) grids, NO data-path collective; the only exchange is the int64[8]
episode-statistics all-reduce after a rollout."""
import os

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of global env ids owned by `rank` (remainder spread over the
    first ranks).  Global ids key the Philox streams and the env->grid map, so results do not
    depend on `world`."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK/WORLD_SIZE/MASTER_*).
    Returns (rank, world, local_rank); a single process needs no process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend)
    return rank, world, local


def all_reduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """In-place SUM of int64 episode counters over all ranks (integer sums: order-independent,
    so N-GPU totals equal the 1-GPU totals bit for bit).  Call it on a per-launch DELTA or on a
    copy of the cumulative local counters -- never twice on the same cumulative buffer."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


class StatsReducer:
    """Global episode statistics of a sharded rollout without stalling the data path.

    After every rollout launch `submit(local_stats)` snapshots this rank's cumulative counters
    (stream-ordered after the kernel) and starts their all-reduce with async_op=True: NCCL runs it
    on its own stream, so the next rollout kernel is not queued behind the collective.  `result()`
    waits for the newest submitted reduction and returns the global cumulative totals."""

    def __init__(self):
        self._buf = None
        self._work = None

    def submit(self, local_stats: torch.Tensor) -> None:
        if self._work is not None:
            self._work.wait()                 # at most one reduction in flight; it overlapped the last kernel
        buf = local_stats.clone()
        self._buf = buf
        self._work = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self._work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=True)

    def result(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = None
        return self._buf


def max_over_ranks(value: float, device=None) -> float:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return value
