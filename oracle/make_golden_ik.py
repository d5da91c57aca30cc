"""Generate tests/golden/ik_cases.npz from the UNMODIFIED reference: TwoJointRobot.inverse_kinematic,
cart_target_valid_check and move_to_joint_pose (robot/two_joint_robot.py:49-113).
Run here:  python oracle/make_golden_ik.py        TEST INFRASTRUCTURE ONLY."""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_boot  # noqa: E402


def main():
    R = ref_boot.boot()
    rng = np.random.default_rng(77)
    n = 3000
    t = rng.uniform(-0.8, 0.8, (n, 2))
    t[:20] = [[0.5, 0.0], [0.0, 0.5], [-0.2, -0.3], [0.7, 0.0], [0.0, -0.7], [0.1, 0.0], [0.0, 0.1], [0.05, 0.05],
              [0.7000001, 0.0], [0.0999, 0.0], [0.3, 0.4], [-0.3, 0.4], [0.42, -0.56], [0.0, 0.0], [0.6, 0.3],
              [-0.7, 0.0], [0.2, 0.2], [-0.1, 0.05], [0.35, -0.1], [0.1000001, 0.0]]
    valid = np.zeros(n, dtype=np.uint8)
    sol = np.zeros((n, 4))
    rob = R.TwoJointRobot(0.0, 0.0)
    with contextlib.redirect_stdout(io.StringIO()):       # "Target out of reach." prints
        for i in range(n):
            s1, s2 = rob.inverse_kinematic(R.Point(t[i, 0], t[i, 1]))
            if s1 is not None:
                valid[i] = 1
                sol[i] = (s1[0], s1[1], s2[0], s2[1])
    m = 500
    start = rng.uniform(-3, 3, (m, 2)); goal = rng.uniform(-3, 3, (m, 2))
    steps = rng.integers(1, 200, m).astype(np.int32)
    end = np.zeros((m, 2))
    for i in range(m):
        rb = R.TwoJointRobot(start[i, 0], start[i, 1])
        rb.move_to_joint_pose(goal[i, 0], goal[i, 1], steps=int(steps[i]))
        end[i] = (rb.joint_1, rb.joint_2)
    out = os.path.join(HERE, "..", "tests", "golden", "ik_cases.npz")
    np.savez_compressed(out, target=t, valid=valid, sol=sol, start=start, goal=goal, steps=steps, end=end)
    print("valid", int(valid.sum()), "of", n, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
