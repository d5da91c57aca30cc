"""world_size-2 gloo test (CPU) of the N>1 host logic: contiguous env slices keyed by GLOBAL env
ids + the int64 statistics all-reduce give the same totals and the same per-env results as one
process owning every env.  The per-shard compute here is the CPU oracle (this is a test)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from abstract_gym_b200.sharding import shard_range
    for n in (0, 1, 7, 4096, 1 << 20, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


N, K, SEED = 3000, 24, 5


def _poses():
    rng = np.random.default_rng(77)
    return rng.uniform(0, 2 * np.pi, N), rng.uniform(0, 2 * np.pi, N)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from abstract_gym_b200.sharding import all_reduce_stats, init_from_env, shard_range
    from oracle import oracle as orc
    init_from_env(backend="gloo")
    lo, hi = shard_range(N, rank, world)
    j1, j2 = _poses()
    st = orc.RolloutState(j1[lo:hi], j2[lo:hi])
    sq, _ = orc.manual_grid()
    rec, stats = orc.rollout(st, K, [sq], env_id0=lo, seed=SEED, threads=1)
    t = torch.from_numpy(stats.copy())
    all_reduce_stats(t)
    # cumulative counters reduced after every launch (bench.py's pattern): no double counting
    from abstract_gym_b200.sharding import StatsReducer
    red, cum = StatsReducer(), torch.zeros(len(stats), dtype=torch.int64)
    for _ in range(3):
        cum += torch.from_numpy(stats)           # "launch" adds this rank's delta to its cumulative counters
        red.submit(cum)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), stats=t.numpy(), j1=st.j1, flags=rec["flags"], lo=lo, hi=hi,
             cum3=red.result().numpy())
    dist.destroy_process_group()


def test_two_rank_rollout_equals_single(tmp_path, oracle):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    j1, j2 = _poses()
    st = oracle.RolloutState(j1, j2)
    sq, _ = oracle.manual_grid()
    rec, stats = oracle.rollout(st, K, [sq], env_id0=0, seed=SEED)
    parts = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(2)]
    for p in parts:
        assert np.array_equal(p["stats"], stats)          # all-reduced totals == single-process totals
        assert np.array_equal(p["cum3"], 3 * stats)
        lo, hi = int(p["lo"]), int(p["hi"])
        assert np.array_equal(p["j1"], st.j1[lo:hi])      # per-env results do not depend on the sharding
        assert np.array_equal(p["flags"], rec["flags"][:, lo:hi])
    assert stats[oracle.ST_ENV_STEPS] == N * K and stats[oracle.ST_EPISODES] > 0
