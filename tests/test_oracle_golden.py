"""Pin the CPU oracle (oracle/ag_oracle.c) to outputs of the UNMODIFIED reference.

Fixtures under tests/golden/ were produced by oracle/make_golden.py from /root/reference.
Everything here is bit-exact (float64 equality), no tolerances.
"""
import hashlib
import json
import math
import os
import struct

import numpy as np
import pytest


@pytest.fixture(scope="module")
def goldens(golden_dir):
    with open(os.path.join(golden_dir, "reference_goldens.json")) as f:
        return json.load(f)


def test_libm_matches_numpy():
    # numpy's float64 sin/cos are glibc's on this image, so the C oracle's
    # FK is bit-identical to the reference's.  Re-checked wherever the tests run.
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(0, 2 * np.pi, 20000), rng.uniform(-100, 100, 20000)])
    assert all(math.sin(v) == s for v, s in zip(x.tolist(), np.sin(x).tolist()))
    assert all(math.cos(v) == c for v, c in zip(x.tolist(), np.cos(x).tolist()))


def test_known_answers(oracle, goldens):
    k = goldens["known"]
    # utils/collision_checker.py:94-96
    hit, aa = oracle.segment_square(0, 0, 1, 2, [0, 0.8, 0.9, 1.4])
    assert hit is True and k["collision_main"] is True and aa == 0
    # TwoJointRobot(0.3, 1.2)
    ex, ey, gx, gy = oracle.forward_kinematics(0.3, 1.2)
    assert [gx, gy] == k["fk_0p3_1p2"]["ee"] == [0.4908419219932445, 0.39781980845470366]
    assert [ex, ey] == k["fk_0p3_1p2"]["elbow"] == [0.38213459565024244, 0.11820808266453582]


def test_manual_grid_corners(oracle):
    # SURVEY.md section 8a G1 [probed] values of environment/occupancy_grid.py:45-50,59-67
    sq, ci = oracle.manual_grid()
    assert sq[0].tolist() == [0.40000000000000013, -0.19999999999999996, 0.6000000000000001, 5.551115123125783e-17]
    assert sq[1].tolist() == [0.6000000000000001, -0.19999999999999996, 0.8, 5.551115123125783e-17]
    assert sq[2].tolist() == [-0.19999999999999996, 0.4, 5.551115123125783e-17, 0.6000000000000001]
    assert ci.tolist() == [51, 52, 21]


def test_grid201(oracle, goldens, golden_dir):
    z = np.load(os.path.join(golden_dir, "grid201_seed0.npz"))
    occ = np.unpackbits(z["occ"])[:201 * 201].reshape(201, 201)
    sq, ci = oracle.grid_squares(occ)
    assert len(sq) == goldens["known"]["grid201_seed0_count"] == 4126
    assert np.array_equal(sq[:64], z["first_squares"]) and np.array_equal(sq[-64:], z["last_squares"])
    assert goldens["known"]["grid201_side"] == 1.6 / 200


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_experiment0_digest(oracle, goldens, idx):
    g = goldens["experiment0"][idx]
    sq, _ = oracle.manual_grid()
    draws = np.random.RandomState(g["seed"]).rand(2 * g["steps"] + 4096)
    rec, resets, final, used = oracle.experiment_loop(g["start_joints"][0], g["start_joints"][1], draws,
                                                      g["steps"], sq)
    h = hashlib.sha256()
    for r in rec:
        h.update(struct.pack("<5d2B", r[0], r[1], r[2], r[3], r[4], int(r[5]), int(r[6])))
    assert h.hexdigest() == g["sha256"]
    assert resets.tolist() == g["resets"]
    assert list(final) == g["final_joints"]
    assert rec[0, :5].tolist() == g["first_record"][:5]


def test_experiment0_survey_digests(goldens):
    # the prefixes quoted in SURVEY.md section 8c
    pre = {0: "61b769515acef880", 1: "3b9d7f40554b423a", 2: "c39bbf257c5247d3"}
    for g in goldens["experiment0"]:
        assert g["sha256"].startswith(pre[g["seed"]])
    assert goldens["experiment0"][0]["resets"] == [3325, 7529, 14999]


def test_predicate_cases(oracle, golden_dir):
    z = np.load(os.path.join(golden_dir, "predicate_cases.npz"))
    seg, sq, abc, signs, out = z["seg"], z["sq"], z["abc"], z["signs"], z["out"]
    lib = oracle.lib()
    import ctypes as C
    n = len(out)
    n_crash = 0
    for i in range(n):
        a, b, c = oracle.line_function(*seg[i])
        assert (a, b, c) == tuple(abc[i]) or (math.isnan(c) and math.isnan(abc[i][2])), i
        v = (C.c_double * 4)()
        s = np.ascontiguousarray(sq[i])
        lib.ago_corner_values(C.c_double(a), C.c_double(b), C.c_double(c), s.ctypes.data_as(C.c_void_p), v)
        assert np.array_equal(np.sign(np.array(list(v))), signs[i]), i
        hit, aa = oracle.segment_square(*seg[i], sq[i])
        if out[i] == 2:        # AttributeError in the reference: defined + counted here
            assert aa == 1, i
            n_crash += 1
        else:
            assert aa == 0 and int(hit) == out[i], i
    assert n_crash > 50 and (out == 1).sum() > 2000 and (out == 0).sum() > 2000


def test_fk_cases(oracle, golden_dir):
    z = np.load(os.path.join(golden_dir, "fk_cases.npz"))
    for j, fk in zip(z["j"], z["fk"]):
        assert oracle.forward_kinematics(j[0], j[1]) == tuple(fk)


GRIDS = ["manual9", "rand9", "rand31", "rand64", "rand6", "matrix5"]


@pytest.mark.parametrize("name", GRIDS)
def test_grid_and_scene_steps(oracle, golden_dir, name):
    z = np.load(os.path.join(golden_dir, "scene_cases.npz"))
    occ, sq_ref, ci_ref = z[name + "/occ"], z[name + "/squares"], z[name + "/cell_index"]
    if name == "manual9":
        sq, ci = oracle.manual_grid()
    else:
        sq, ci = oracle.grid_squares(occ, float(z[name + "/env_size"]))
    assert np.array_equal(sq, sq_ref) and np.array_equal(ci, ci_ref)
    j0, acts, out, fh, ee = (z[name + "/" + k] for k in ("j0", "actions", "out", "first_hit", "ee"))
    n_env, n_steps = acts.shape[:2]
    j1, j2 = j0[:, 0].copy(), j0[:, 1].copy()
    reward = np.zeros(n_env); flags = np.zeros(n_env, dtype=np.uint8)
    for t in range(n_steps):
        r = oracle.step_batch(j1, j2, acts[:, t], reward, flags, sq, ci)
        assert np.array_equal(j1, out[:, t, 0]) and np.array_equal(j2, out[:, t, 1])
        assert np.array_equal(reward, out[:, t, 2])
        assert np.array_equal((flags & 2) != 0, out[:, t, 3] != 0)
        assert np.array_equal((flags & 1) != 0, out[:, t, 4] != 0)
        assert np.array_equal(r["first_hit"], fh[:, t])
        assert np.array_equal(r["ee"], ee[:, t])
        assert r["axis_aligned"] == 0
    assert (out[:, :, 4] != 0).any()


def test_reach_sequences(oracle, golden_dir):
    z = np.load(os.path.join(golden_dir, "scene_cases.npz"))
    start, seq = z["reach/start"], z["reach/seq"]
    sq, ci = oracle.manual_grid()
    n = len(start)
    j1, j2 = start[:, 0].copy(), start[:, 1].copy()
    reward = np.zeros(n); flags = np.zeros(n, dtype=np.uint8)
    for t in range(seq.shape[1]):
        oracle.step_batch(j1, j2, seq[:, t, 0:2], reward, flags, sq, ci)
        assert np.array_equal(j1, seq[:, t, 2]) and np.array_equal(j2, seq[:, t, 3])
        assert np.array_equal(reward, seq[:, t, 4])
        assert np.array_equal((flags & 2) != 0, seq[:, t, 5] != 0)
        assert np.array_equal((flags & 1) != 0, seq[:, t, 6] != 0)
    assert (seq[:, :, 5] != 0).any(), "fixture must exercise done=True"


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32 10 rounds
    assert oracle.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert oracle.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == \
        (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert oracle.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    u = oracle.philox_uniform2(7, 123456789012, 5, 1)
    assert 0.0 <= u[0] < 1.0 and 0.0 <= u[1] < 1.0 and u[0] != u[1]


def test_ik_and_move_to_joint_pose_goldens(oracle, golden_dir):
    """robot/two_joint_robot.py:49-113 vs fixtures from the live reference (oracle/make_golden_ik.py).
    move_to_joint_pose and the reachability test are bit-exact; the IK angles agree to 4 ulp of pi: numpy's
    float64 arccos is its own SIMD implementation and differs from glibc's acos by 1 ulp on 9 % of arguments
    (measured here), so the C restatement cannot be bit-identical off the hot path."""
    z = np.load(os.path.join(golden_dir, "ik_cases.npz"))
    t, valid, sol = z["target"], z["valid"], z["sol"]
    assert 1000 < int(valid.sum()) < len(valid)
    for i in range(len(t)):
        ok, s = oracle.inverse_kinematics(float(t[i, 0]), float(t[i, 1]))
        assert ok == bool(valid[i]), (i, t[i])
        if ok:
            assert np.max(np.abs(np.array(s) - sol[i])) <= 4 * 4.45e-16, (i, t[i], s, sol[i])
    ok, s = oracle.inverse_kinematics(0.5, 0.0)                 # the reference's __main__ known answer
    assert ok and np.allclose(s, [-0.64350111, 0.92729522, 0.64350111, -0.92729522], atol=5e-9)
    ok, s = oracle.inverse_kinematics(-0.2, -0.3, corrected=True)
    ex, ey, gx, gy = oracle.forward_kinematics(s[0], s[1])
    assert abs(gx + 0.2) < 1e-12 and abs(gy + 0.3) < 1e-12      # atan2 variant reaches targets with y < 0
    for i in range(len(z["steps"])):
        e1, e2 = oracle.move_to_joint_pose(z["start"][i, 0], z["start"][i, 1], z["goal"][i, 0], z["goal"][i, 1],
                                           int(z["steps"][i]))
        assert [e1, e2] == z["end"][i].tolist()
