"""Time the UNMODIFIED reference's own loop body (experiment/experiment_0.py:20-34) in this process.

    python oracle/ref_loop.py --steps 20000 --seed 0     ->  one JSON line {"steps": .., "seconds": .., "resets": ..}

MEASUREMENT INFRASTRUCTURE ONLY (bench.py's cpu_baseline leg runs one of these per host core).  Boots the reference from
the vendored copy (oracle/vendor_reference.py) or /root/reference through oracle/ref_boot.py.
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    import vendor_reference
    root = vendor_reference.vendored_root()
    if root:
        os.environ["ABSTRACT_GYM_REFERENCE"] = root
    import ref_boot
    import numpy as np
    R = ref_boot.boot()
    np.random.seed(a.seed)
    rob = R.TwoJointRobot(joint_1=1.0, joint_2=2.5)                  # experiment_0.py:13-16
    occ = R.OccupancyGrid(size=9, random_obstacle=False)
    s = R.Scene(rob, occ)
    s.random_valid_pose()
    resets = 0
    t0 = time.perf_counter()
    for _ in range(a.steps):                                         # experiment_0.py:20-34
        act = s.sample_action(scale_factor=0.1)
        j1, j2, r, d, c = s.step(act)
        if d or c:
            resets += 1
            s.reset()
    dt = time.perf_counter() - t0
    print(json.dumps({"steps": a.steps, "seconds": dt, "resets": resets}))


if __name__ == "__main__":
    main()
