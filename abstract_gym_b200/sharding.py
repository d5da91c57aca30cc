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


def gpu_cpu_affinity(index: int):
    """CPU set NVML reports as local to GPU `index` (the 'CPU Affinity' column of `nvidia-smi topo -m`), or None."""
    import re
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                             timeout=20).stdout
    except Exception:
        return None
    for line in out.splitlines():
        tok = line.split()
        if not tok or tok[0] != "GPU%d" % index:
            continue
        for t in tok[1:]:
            if re.fullmatch(r"\d+(-\d+)?(,\d+(-\d+)?)*", t) and ("-" in t or "," in t):
                cpus = set()
                for part in t.split(","):
                    lo, _, hi = part.partition("-")
                    cpus.update(range(int(lo), int(hi or lo) + 1))
                return cpus
    return None


def bind_to_gpu_numa(index: int) -> bool:
    """Pin this process to the CPUs next to its GPU before it allocates pinned host buffers: page-locked memory is
    placed on the NUMA node of the allocating thread, and with 8 ranks streaming 1.4 GB per launch each, copies
    that cross the socket interconnect share its bandwidth."""
    cpus = gpu_cpu_affinity(index)
    if not cpus:
        return False
    try:
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if target:
            os.sched_setaffinity(0, target)
            return True
    except Exception:
        pass
    return False


def nccl_options():
    """NCCL on a high-priority stream: the statistics all-reduce is a tiny kernel that must slip in between
    the blocks of a rollout kernel that fills every SM, not wait for it to drain."""
    try:
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        return opts
    except Exception:
        return None


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
            dist.init_process_group(backend=backend, pg_options=nccl_options(), device_id=torch.device("cuda", local))
        else:
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
    on its own stream, so the next rollout kernel is not queued behind the collective.  Up to
    `depth` reductions stay in flight (the collective's tiny kernel may only get an SM slot when
    the rollout kernel that was launched right behind it drains).  `result()` waits for the newest
    submitted reduction and returns the global cumulative totals."""

    def __init__(self, depth: int = 2):
        self._depth = max(1, int(depth))
        self._pending = []          # [(buffer, work)] oldest first

    def submit(self, local_stats: torch.Tensor) -> None:
        while len(self._pending) >= self._depth:
            _, work = self._pending.pop(0)
            if work is not None:
                work.wait()         # stream-level wait; it had `depth` launches to finish
        buf = local_stats.clone()
        work = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=True)
        self._pending.append((buf, work))

    def result(self) -> torch.Tensor:
        if not self._pending:
            return None
        for _, work in self._pending:
            if work is not None:
                work.wait()
        buf = self._pending[-1][0]
        self._pending = [(buf, None)]
        return buf


def max_over_ranks(value: float, device=None) -> float:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return value
