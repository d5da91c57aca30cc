"""Generate tests/golden/data_list_*.txt: the reference's own data product (experiment/experiment_0.py:54-57,
one `"%s\\n" % episode` line per completed episode) from the UNMODIFIED reference, single env, seeded.

Run here (the container that has /root/reference):  python oracle/make_golden_datalist.py
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_boot  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def thread_function(R, seed, steps, occ_matrix):
    """experiment/experiment_0.py:11-37 with a seeded global numpy stream and an explicit start pose
    (the reference's TwoJointRobot() default is an import-time random)."""
    np.random.seed(seed)
    rob = R.TwoJointRobot(joint_1=1.0, joint_2=2.5)
    occ = R.OccupancyGrid(size=9, random_obstacle=False, obstacle_probability=0.01)
    if occ_matrix is not None:
        occ.load_from_matrix(occ_matrix)
    s = R.Scene(robot=rob, env=occ, visualize=False)
    s.random_valid_pose()
    record, record_list = [], []
    for i in range(steps):
        action = s.sample_action(scale_factor=0.1)
        j1, j2, step_reward, done, collision = s.step(action)
        record.append([j1, j2, action[0], action[1], step_reward, done, collision])
        if done or collision:
            record_list.append(record.copy())
            record.clear()
            s.reset()
    return record_list


# a denser map than scene_0's, so that a few hundred steps hold several episodes (small fixture)
DENSE = np.array([[0, 0, 0, 1, 0, 0, 0, 0, 0],
                  [0, 1, 0, 0, 0, 0, 1, 0, 0],
                  [0, 0, 0, 1, 0, 0, 0, 0, 1],
                  [1, 0, 0, 0, 0, 1, 0, 0, 0],
                  [0, 0, 0, 0, 0, 0, 0, 1, 0],
                  [0, 0, 1, 0, 0, 0, 1, 1, 0],
                  [0, 0, 0, 0, 1, 0, 0, 0, 0],
                  [0, 1, 0, 0, 0, 0, 0, 0, 0],
                  [0, 0, 0, 0, 0, 1, 0, 0, 1]])


def main():
    R = ref_boot.boot()
    for name, seed, steps, occ in (("dense9_seed3", 3, 400, DENSE), ("dense9_seed4", 4, 400, DENSE)):
        eps = thread_function(R, seed, steps, occ)
        path = os.path.join(OUT, "data_list_%s.txt" % name)
        with open(path, "w") as f:
            for d in eps:                       # experiment_0.py:54-57
                f.write("%s\n" % d)
        print(name, "episodes", len(eps), "lengths", [len(e) for e in eps], os.path.getsize(path), "bytes")
    np.save(os.path.join(OUT, "data_list_dense9_occ.npy"), DENSE.astype(np.uint8))


if __name__ == "__main__":
    main()
