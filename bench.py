#!/usr/bin/env python
"""bench.py -- env-steps/s of the scene_0 step/reset hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU arithmetic (oracle port)

One bench "step" = ONE fused rollout launch (K4) over this rank's environments:
2^20 envs x 64 env-steps, RECORD mode (reads float32 actions [64,N,2], writes the trajectory
records the reference appends per step: joint_1, joint_2, reward, flags) -- BASELINE.json
configs[2] ("scene_0 with 1M envs per GPU, fused 64-step rollout kernel, at 1/2/4/8 B200").
N>1: one process per GPU (torchrun), contiguous env slices with global env ids, replicated grid,
no data-path collective; the int64[8] episode statistics are all-reduced (NCCL) after every launch,
inside the timed region.  `value` = env-steps of all ranks / max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG_RECORD = 22.0   # algorithmic bytes per env-step, RECORD mode (SURVEY.md 8d / DESIGN.md): 8 read + 13 written + ~1 amortised state
METRIC = "env-steps/sec (FK+collision+reward), scene_0 fused 64-step rollout"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=1 << 20, help="environments per GPU")
    ap.add_argument("--rollout-steps", type=int, default=64, help="env-steps fused per launch (K)")
    ap.add_argument("--engine", default="fast", choices=["fast", "exact"])
    ap.add_argument("--mode", default="record", choices=["record", "stats"])
    ap.add_argument("--grid", default="scene0", choices=["scene0", "c4", "c5"],
                    help="scene0: manual 9x9 map (BASELINE configs[2], the default bench line); c4: one 1024x1024 "
                         "Bernoulli(0.002) map (configs[3]); c5: a distinct 256x256 Bernoulli(0.008) map per 256 envs (configs[4])")
    ap.add_argument("--chunk-envs", type=int, default=1 << 17, help="envs per pipeline stage of the host-buffer (e2e) path")
    ap.add_argument("--chunk-steps", type=int, default=4, help="steps per pipeline stage of the e2e path (0: slice over envs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-envs", type=int, default=1 << 17)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic():
    """per-launch DRAM bytes of the rollout kernel from the committed ncu --set full capture"""
    path = os.path.join(ROOT, "profiles", "rollout_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            pass
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], 0.0, set(), 0.0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for r in self.rows if t0 <= r[0] <= t1 + 0.1] or self.rows[-3:]
        for _, line in rows:
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1])); power = max(power, float(f[2]))
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "power_w_max": power, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_reference(args, rank, world):
    """The reference's own CPU arithmetic for the path, on the host cores.  The reference is pure
    Python and cannot travel to the GPU box, so this arm runs its bit-exact C restatement (oracle/,
    pinned to the reference's goldens) with all host threads -- a faster CPU baseline than the
    reference's interpreter loop (measured here: 9.5e3 env-steps/s/core)."""
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as orc
    K = args.rollout_steps
    rng = np.random.default_rng(0)
    sqs, epg, n = oracle_workload(args, orc, np)
    st = orc.RolloutState(rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 2 * np.pi, n))
    actions = ((rng.random((K, n, 2)) - 0.5) * 0.1).astype(np.float32) if args.mode == "record" else None
    cores = host_cores()                             # torchrun exports OMP_NUM_THREADS=1: ask for every core explicitly
    for _ in range(args.warmup):
        orc.rollout(st, K, sqs, envs_per_grid=epg, seed=0, actions_f32=actions, record=args.mode == "record", threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.rollout(st, K, sqs, envs_per_grid=epg, seed=0, actions_f32=actions, record=args.mode == "record", threads=cores)
    dt = time.perf_counter() - t0
    value = n * K * args.steps / dt
    sample = "%d envs x %d env-steps per step (1/%d of one GPU's batch), OpenMP over envs" % (n, K, max(1, args.envs // n))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def oracle_workload(args, orc, np):
    """(squares per grid, envs_per_grid, sample envs) of the CPU arms for --grid.  The oracle tests every obstacle
    (the reference's O(#obstacles) loop), so the high-resolution maps get a smaller sample."""
    if args.grid == "scene0":
        return [orc.manual_grid()[0]], None, args.cpu_sample_envs
    if args.grid == "c4":
        occ = (np.random.default_rng(4).random((1024, 1024)) < 0.002).astype(np.uint8)
        return [orc.grid_squares(occ)[0]], None, min(args.cpu_sample_envs, 1 << 11)
    rng = np.random.default_rng(5)
    occs = [(rng.random((256, 256)) < 0.008).astype(np.uint8) for _ in range(8)]
    return [orc.grid_squares(o)[0] for o in occs], 256, min(args.cpu_sample_envs, 1 << 11)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


GRID_NAMES = {"scene0": "scene_0 manual 9x9 grid (3 obstacles)",
              "c4": "one 1024x1024 bit-packed grid, Bernoulli(0.002) obstacles (128 KiB, L2-resident)",
              "c5": "a distinct 256x256 Bernoulli(0.008) grid per batch of 256 envs (8 KiB each, staged by TMA)"}


def workload_config(args, world):
    return {"workload": "%s, %d envs/GPU, fused %d-step rollout (K4), %s mode, "
                        "random actions (u-0.5)*0.1, auto-reset" % (GRID_NAMES[args.grid], args.envs, args.rollout_steps,
                                                                    args.mode.upper()),
            "envs_per_gpu": args.envs, "rollout_steps": args.rollout_steps, "mode": args.mode, "engine": args.engine,
            "arithmetic": ("float64 decisions: float32 interval filter, float64 filter, then the reference's float64 "
                           "operation order for what they cannot settle") if args.engine == "fast" else "float64",
            "grid": args.grid, "parallelism": "env-sharded x%d, replicated grid, stats all-reduce" % world,
            "l2_policy": "inputs_exceed_l2 (%.0f MB streamed per launch)" %
                         (args.envs * args.rollout_steps * (21 if args.mode == "record" else 0) / 1e6)}


def cpu_baseline(args):
    import numpy as np
    from oracle import oracle as orc
    K = args.rollout_steps
    rng = np.random.default_rng(0)
    sqs, epg, n = oracle_workload(args, orc, np)
    n *= 2
    st = orc.RolloutState(rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 2 * np.pi, n))
    actions = ((rng.random((K, n, 2)) - 0.5) * 0.1).astype(np.float32) if args.mode == "record" else None
    cores = host_cores()
    orc.rollout(st, K, sqs, envs_per_grid=epg, seed=0, actions_f32=actions, record=args.mode == "record", threads=cores)
    reps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 10.0:
        orc.rollout(st, K, sqs, envs_per_grid=epg, seed=0, actions_f32=actions, record=args.mode == "record", threads=cores)
        reps += 1
    dt = time.perf_counter() - t0
    return {"value": n * K * reps / dt, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": "%d x (%d envs x %d env-steps), %.1f s, OpenMP over envs" % (reps, n, K, dt)}


def run_ours(args, rank, world, local):
    import numpy as np
    import torch
    import torch.distributed as dist
    import abstract_gym_b200 as ag

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    from abstract_gym_b200.sharding import bind_to_gpu_numa
    numa_bound = bind_to_gpu_numa(local) if (world > 1 and not os.environ.get("AG_NO_NUMA_BIND")) else False
    n, K = args.envs, args.rollout_steps
    record = args.mode == "record"
    lo = rank * n                                    # weak scaling: every rank owns n envs, global ids [lo, lo+n)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    if args.grid == "scene0":
        grid = ag.OccupancyGrid(size=9, random_obstacle=False)
    elif args.grid == "c4":
        grid = ag.OccupancyGrid(size=9, random_obstacle=False)
        grid.load_from_matrix((np.random.default_rng(4).random((1024, 1024)) < 0.002).astype(np.uint8))
    else:
        ggen = torch.Generator(device=dev).manual_seed(5)     # same maps on every rank; envs pick them by global id
        grid = ag.BatchedOccupancyGrid.random(max(1, world * n // 256), 256, 0.008, 256, device=dev, generator=ggen,
                                              clear_base_cells=2)
    robot = ag.BatchedTwoJointRobot.random(n, device=dev, generator=gen)
    scene = ag.BatchedScene(robot, grid, engine=args.engine, seed=0, env_id0=lo)
    scene.random_valid_pose()                        # experiment_0.py:16
    actions = None
    rec = None
    if record:
        actions = ((torch.rand(K, n, 2, device=dev, generator=gen) - 0.5) * 0.1).to(torch.float32)
        rec = scene.alloc_records(K)

    def one_step():
        scene.rollout(K, actions=actions, record=record, out=rec)
        scene.all_reduce_stats(wait=False)           # async: overlaps the next launch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local) if rank == 0 else None     # covers warm-up, the timed launches and the e2e loop
    wall_load0 = time.time()
    for _ in range(args.warmup):
        one_step()
    barrier()
    launches0 = ag.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    start.record()
    for a, b in evs:
        a.record()                                   # same stream the kernel is launched on (torch current stream)
        scene.rollout(K, actions=actions, record=record, out=rec)
        b.record()
        scene.all_reduce_stats(wait=False)
    loop_end = torch.cuda.Event(enable_timing=True)
    loop_end.record()
    scene.global_stats()                             # the last reduction is inside the timed region
    stop.record()
    barrier()
    wall1 = time.time()
    total_ms = start.elapsed_time(stop)
    kern_ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    tail_ms = loop_end.elapsed_time(stop)            # waiting for the last statistics reductions
    launches = ag.launch_count() - launches0
    t = torch.tensor([total_ms, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms = t.tolist()
    value = world * n * K * args.steps / (total_ms * 1e-3)

    # ---- end to end: host actions -> H2D -> K4 -> D2H records, through the public API ------------
    e2e = None
    if not args.no_e2e:
        hact = None
        hout = None
        if record:
            hact = torch.empty(K, n, 2, dtype=torch.float32, pin_memory=True)
            hact.copy_(actions)
            hout = scene.alloc_records(K, pinned_host=True)
        reps = max(2, min(args.steps, 5))
        scene.rollout_host(K, hact, hout, chunk_envs=args.chunk_envs, chunk_steps=args.chunk_steps)   # warm-up (allocates the staging pipeline)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            scene.rollout_host(K, hact, hout, chunk_envs=args.chunk_envs, chunk_steps=args.chunk_steps)   # returns after records + stats are in host memory
            scene.all_reduce_stats(wait=False)
        scene.global_stats()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * n * K * reps / dt, "unit": "env-steps/s",
               "h2d_bytes_per_step": (K * n * 8) if record else 0,
               "d2h_bytes_per_step": (K * n * 13 if record else 0) + 8 * 8 * 3,
               "reps": reps, "ms_per_step": 1e3 * dt / reps,
               "api": "BatchedScene.rollout_host -> ag_rollout_host (pinned host buffers, 3-stream pipeline over %s)"
                      % ("slices of %d steps" % args.chunk_steps if args.chunk_steps > 0 else "slices of %d envs" % args.chunk_envs)}
        if record:   # same call with compact records (no reward plane: it is a function of the flags), reported beside it
            hout_c = {k: v for k, v in hout.items() if k != "reward"}
            scene.rollout_host(K, hact, hout_c, chunk_envs=args.chunk_envs, chunk_steps=args.chunk_steps)
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                scene.rollout_host(K, hact, hout_c, chunk_envs=args.chunk_envs, chunk_steps=args.chunk_steps)
            barrier()
            dtc = time.perf_counter() - t0
            ttc = torch.tensor([dtc], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ttc, op=dist.ReduceOp.MAX)
            e2e["compact_records"] = {"value": world * n * K * reps / float(ttc.item()), "unit": "env-steps/s",
                                      "d2h_bytes_per_step": K * n * 9 + 8 * 8 * 3,
                                      "note": "reward plane not copied: reward = f(flags) in a rollout record"}
    clocks = sampler.stop(wall_load0, time.time()) if sampler else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed launches + e2e loop (the timed launches alone last %.1f ms)" % total_ms
    stats = dict(zip(ag.STAT_NAMES, scene.all_reduce_stats().tolist()))   # global totals over all ranks
    if rank != 0:
        return
    peak, peak_src = peaks()
    b_alg = B_ALG_RECORD if record else 64.0 / K
    achieved = n * K * b_alg / (kern_ms * 1e-3) / 1e9
    traffic = measured_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, world),
        "e2e": e2e, "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None if traffic is None else traffic.get("dram_bytes_per_launch"),
                     "peak_source": peak_src, "bytes_per_env_step": b_alg, "kernel_ms": kern_ms,
                     "kernel": "k_rollout", "note": "compute-bound path: see DESIGN.md roofline section"},
        "clocks": clocks, "episode_stats": stats, "filter_diag_rank0": scene.diag_dict(), "numa_bound": numa_bound,
        "stats_reduce_tail_ms_rank0": tail_ms,
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        from abstract_gym_b200.sharding import nccl_options
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local), pg_options=nccl_options())
    run_ours(args, rank, world, local)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
