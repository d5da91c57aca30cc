"""Shard invariance on a real device: two processes, each owning a contiguous slice of the global env
ids on the GPU (the test box has one B200, so both ranks share it; the rendezvous and the statistics all-reduce go
through gloo), must reproduce the one-process run bit for bit: per-env records and final state, and the all-reduced
episode statistics equal the single-GPU totals."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, K, SEED = 1 << 15, 32, 11


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene(ag, torch, kind, lo, hi):
    rng = np.random.default_rng(123)
    j1, j2 = rng.uniform(0, 2 * np.pi, N), rng.uniform(0, 2 * np.pi, N)
    if kind == "scene0":
        grid = ag.OccupancyGrid(size=9, random_obstacle=False)
    else:                                   # per-batch maps: every rank holds all maps, envs pick theirs by GLOBAL id
        ggen = torch.Generator(device="cuda").manual_seed(5)
        grid = ag.BatchedOccupancyGrid.random(N // 256, 256, 0.008, 256, device="cuda", generator=ggen, clear_base_cells=2)
    rb = ag.BatchedTwoJointRobot(torch.as_tensor(j1[lo:hi], device="cuda"), torch.as_tensor(j2[lo:hi], device="cuda"))
    sc = ag.BatchedScene(rb, grid, engine="fast", seed=SEED, env_id0=lo)
    sc.random_valid_pose()
    return sc


def _worker(rank, world, port, out_dir, kind):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import abstract_gym_b200 as ag
    from abstract_gym_b200.sharding import shard_range
    dist.init_process_group(backend="gloo")
    torch.cuda.set_device(0)
    lo, hi = shard_range(N, rank, world)
    sc = _scene(ag, torch, kind, lo, hi)
    rec = None
    for _ in range(2):                      # two launches: the draw counters and sticky state carry over
        rec = sc.rollout(K)
    local = sc.stats.cpu().clone()
    dist.all_reduce(local, op=dist.ReduceOp.SUM)      # the int64[8] statistics all-reduce (gloo here, NCCL in bench.py)
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, "%s_r%d.npz" % (kind, rank)), stats=local.numpy(), j1=sc.robot.joint_1.cpu().numpy(),
             flags=rec["flags"].cpu().numpy(), rj2=rec["j2"].cpu().numpy(), lo=lo, hi=hi)
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["scene0", "c5"])
def test_two_gpu_ranks_equal_one(tmp_path, kind):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import torch.multiprocessing as mp
    import abstract_gym_b200 as ag
    port = _free_port()
    mp.get_context("spawn")
    mp.spawn(_worker, args=(2, port, str(tmp_path), kind), nprocs=2, join=True)
    sc = _scene(ag, torch, kind, 0, N)
    rec = None
    for _ in range(2):
        rec = sc.rollout(K)
    torch.cuda.synchronize()
    total = sc.stats.cpu().numpy()
    flags, rj2, j1 = rec["flags"].cpu().numpy(), rec["j2"].cpu().numpy(), sc.robot.joint_1.cpu().numpy()
    for r in range(2):
        p = np.load(os.path.join(str(tmp_path), "%s_r%d.npz" % (kind, r)))
        lo, hi = int(p["lo"]), int(p["hi"])
        assert np.array_equal(p["stats"], total)                   # all-reduced totals == the one-GPU totals
        assert np.array_equal(p["j1"], j1[lo:hi])                  # per-env results do not depend on the sharding
        assert np.array_equal(p["flags"], flags[:, lo:hi]) and np.array_equal(p["rj2"], rj2[:, lo:hi])
    assert total[3] == 2 * N * K and total[0] > 0
