"""CPU-only checks of the C-ABI boundary: the library loads, exports every symbol the header
declares, and its host-only helpers agree with the reference goldens.  No device compute here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from abstract_gym_b200 import _lib
    return _lib.load()


def test_header_symbols_exported(lib):
    from abstract_gym_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "abstract_gym_b200.h")).read()
    declared = set(re.findall(r"AG_API[^;(]*?\b(ag_\w+)\s*\(", hdr))
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    assert declared == set(_lib.SYMBOLS), "ctypes table and header disagree: %s" % (declared ^ set(_lib.SYMBOLS))


def test_struct_layouts_match_header():
    from abstract_gym_b200 import _lib
    assert C.sizeof(_lib.Params) == 11 * 8 + 2 * 4
    assert C.sizeof(_lib.Grid) == 3 * 8 + 2 * 8 + 4 * 4 + 2 * 8 + 8 + 8
    assert C.sizeof(_lib.RolloutArgs) == 8 + 8 + 4 + 4 + 8 + 8 + 8 + 4 + 4 + 14 * 8 + 8 + 8 + 8 + 4 + 4


def test_default_params_are_the_reference_literals(lib):
    from abstract_gym_b200 import _lib
    p = _lib.default_params()
    assert (p.link_1, p.link_2) == (0.4, 0.3)                 # robot/two_joint_robot.py:12-13
    assert (p.target_x, p.target_y) == (-0.2, -0.3)           # scenario/scene_0.py:17
    assert (p.target_j1, p.target_j2) == (1.1, -0.2)          # scenario/scene_0.py:30
    assert p.reach_eps == 2e-3 and p.section_eps == 1e-10
    assert (p.reward_collision, p.reward_reach) == (-1e3, 1e4)
    assert p.action_scale == 0.1 and p.choose_j_tar == 0


def test_status_strings(lib):
    from abstract_gym_b200 import _lib
    assert _lib.status_string(0) == "ok"
    assert _lib.status_string(-5) == "The matrix is not square."   # occupancy_grid.py:81
    assert "NULL" in _lib.status_string(-1)


@pytest.mark.parametrize("S", [2, 5, 9, 31, 32, 33, 64, 201, 256, 1024])
def test_grid_tables_match_oracle(lib, oracle, S):
    from abstract_gym_b200 import _lib
    spad = (S + 1) & ~1
    mx, my = np.zeros(spad), np.zeros(spad)
    side = C.c_double()
    assert lib.ag_grid_tables_host(S, 1.6, mx.ctypes.data_as(C.c_void_p), my.ctypes.data_as(C.c_void_p), C.byref(side)) == 0
    occ = np.eye(S, dtype=np.uint8)          # obstacle (r=c=i): min_x = table_x[i], min_y = table_y[i]
    sq, _ = oracle.grid_squares(occ)
    assert np.array_equal(sq[:, 0], mx[:S]) and np.array_equal(sq[:, 1], my[:S])
    assert np.array_equal(sq[:, 2], mx[:S] + side.value) and np.array_equal(sq[:, 3], my[:S] + side.value)
    assert side.value == 1.6 / (S - 1)
    assert lib.ag_grid_words_per_row(S) == (S + 31) // 32
    assert lib.ag_grid_stride_words(S) % 4 == 0 and lib.ag_grid_stride_words(S) >= S * ((S + 31) // 32)


def test_grid_tables_match_reference_goldens(lib, golden_dir):
    z = np.load(os.path.join(golden_dir, "scene_cases.npz"))
    for name in ("manual9", "rand9", "rand31", "rand64", "rand6", "matrix5"):
        occ, sq = z[name + "/occ"], z[name + "/squares"]
        S = occ.shape[0]
        spad = (S + 1) & ~1
        mx, my = np.zeros(spad), np.zeros(spad)
        side = C.c_double()
        lib.ag_grid_tables_host(S, float(z[name + "/env_size"]), mx.ctypes.data_as(C.c_void_p),
                                my.ctypes.data_as(C.c_void_p), C.byref(side))
        ci = z[name + "/cell_index"]
        r, c = ci // S, ci % S
        assert np.array_equal(sq[:, 0], mx[c]) and np.array_equal(sq[:, 1], my[r])
        assert np.array_equal(sq[:, 2], mx[c] + side.value) and np.array_equal(sq[:, 3], my[r] + side.value)


def test_grid_pack_host(lib):
    rng = np.random.default_rng(3)
    for S in (5, 9, 32, 33, 70):
        occ = (rng.random((S, S)) < 0.3).astype(np.uint8)
        words = np.zeros(lib.ag_grid_stride_words(S), dtype=np.uint32)
        assert lib.ag_grid_pack_host(occ.ctypes.data_as(C.c_void_p), S, S, words.ctypes.data_as(C.c_void_p)) == 0
        wpr = (S + 31) // 32
        for r in range(S):
            for c in range(S):
                assert ((words[r * wpr + c // 32] >> (c % 32)) & 1) == occ[r, c]
    occ = np.zeros((4, 5), dtype=np.uint8)
    words = np.zeros(64, dtype=np.uint32)
    assert lib.ag_grid_pack_host(occ.ctypes.data_as(C.c_void_p), 4, 5, words.ctypes.data_as(C.c_void_p)) == -5
    assert lib.ag_grid_pack_host(None, 4, 4, words.ctypes.data_as(C.c_void_p)) == -1


def test_argument_validation_without_device(lib):
    """bad arguments are rejected before anything touches the device"""
    from abstract_gym_b200 import _lib
    p = _lib.default_params()
    g = _lib.Grid()
    assert lib.ag_collision_check(p, g, None, None, None, None, 4, 0, 0, None) == -1      # NULL grid pointers
    g.bits = g.min_x = g.min_y = 256
    g.S, g.words_per_row, g.n_grids, g.grid_stride_words, g.envs_per_grid = 9, 1, 1, 12, 1
    g.side, g.env_size = 0.2, 1.6
    assert lib.ag_collision_check(p, g, None, None, None, None, -1, 0, 0, None) == -2     # n < 0
    assert lib.ag_collision_check(p, g, None, None, None, None, 0, 0, 0, None) == 0       # empty batch is fine
    assert lib.ag_collision_check(p, g, None, None, None, None, 4, 0, 0, None) == -1      # NULL arrays
    g.words_per_row = 2
    assert lib.ag_collision_check(p, g, None, None, None, None, 0, 0, 0, None) == -2      # inconsistent grid
    g.words_per_row = 1
    a = _lib.RolloutArgs()
    a.n, a.K = 4, 0
    assert lib.ag_rollout(p, g, C.byref(a), None) == -2                                   # K < 1
    a.K = 8
    assert lib.ag_rollout(p, g, C.byref(a), None) == -1                                   # NULL state
    assert lib.ag_rollout(p, g, None, None) == -1
    assert lib.ag_launch_count() == 0


def test_product_does_not_touch_the_oracle():
    """the oracle is test infrastructure: nothing under abstract_gym_b200/ may reference it"""
    pkg = os.path.join(ROOT, "abstract_gym_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "__none__", "%s mentions the oracle" % f


def test_no_device_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import abstract_gym_b200 as ag
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ag.BatchedTwoJointRobot([0.1], [0.2])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ag.TwoJointRobot(0.1, 0.2).end_effector()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ag.OccupancyGrid(size=9, random_obstacle=False).device_grid()


def test_reference_host_api_shapes(golden_dir):
    """OccupancyGrid keeps the reference's attributes, order and corner values (host, setup-time)."""
    import abstract_gym_b200 as ag
    z = np.load(os.path.join(golden_dir, "scene_cases.npz"))
    g = ag.OccupancyGrid(size=9, random_obstacle=False)
    occ, coord, obs, side = g.get_occupancy_grid()
    got = np.array([[o.min_x, o.min_y, o.max_x, o.max_y] for o in obs])
    assert np.array_equal(got, z["manual9/squares"]) and side == 0.2 and occ.shape == (9, 9)
    assert coord[0].tolist() == [0.40000000000000013, -0.19999999999999996]
    for name, S, p, seed in (("rand9", 9, 0.1, 11), ("rand31", 31, 0.01, 12), ("rand64", 64, 0.02, 13)):
        np.random.seed(seed)
        g = ag.OccupancyGrid(size=S, random_obstacle=True, obstacle_probability=p)
        got = np.array([[o.min_x, o.min_y, o.max_x, o.max_y] for o in g.obstacle_list])
        assert np.array_equal(got, z[name + "/squares"]) and np.array_equal(g.occ != 0, z[name + "/occ"] != 0)
    mat = np.array([[0, 0, 1, 0, 0], [0, 1, 1, 0, 0], [0, 1, 1, 0, 0], [0, 0, 0, 0, 0], [0, 1, 0, 0, 1]])
    g.load_from_matrix(mat)
    got = np.array([[o.min_x, o.min_y, o.max_x, o.max_y] for o in g.obstacle_list])
    assert np.array_equal(got, z["matrix5/squares"]) and g.size == 5
    g.load_from_matrix(np.zeros((3, 4)))      # prints "The matrix is not square." and leaves the grid alone
    assert g.size == 5
    g.load_from_matrix(np.zeros((4, 4)))      # empty grid accepted (the reference raises ValueError)
    assert g.obstacle_list == [] and g.size == 4


def test_ik_known_answer(golden_dir):
    import json
    import abstract_gym_b200 as ag
    k = json.load(open(os.path.join(golden_dir, "reference_goldens.json")))["known"]
    s1, s2 = ag.TwoJointRobot(0.0, 0.0).inverse_kinematic(ag.Point(0.5, 0.0))   # two_joint_robot.py:115-120
    assert s1.tolist() == k["ik_s1"] and s2.tolist() == k["ik_s2"]
