"""ctypes binding of the CPU oracle (oracle/ag_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package (abstract_gym_b200/) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libag_oracle.so")

ST_EPISODES, ST_COLLISIONS, ST_SUCCESSES, ST_ENV_STEPS, ST_EP_LEN_SUM, ST_RETURN_MILLI, \
    ST_STUCK_RESETS, ST_AXIS_ALIGNED, ST_COUNT = range(9)
FLAG_COLLISION, FLAG_DONE = 1, 2


class Params(C.Structure):
    _fields_ = [("link_1", C.c_double), ("link_2", C.c_double),
                ("target_x", C.c_double), ("target_y", C.c_double),
                ("target_j1", C.c_double), ("target_j2", C.c_double),
                ("reach_eps", C.c_double), ("section_eps", C.c_double),
                ("reward_collision", C.c_double), ("reward_reach", C.c_double),
                ("action_scale", C.c_double),
                ("choose_j_tar", C.c_int32), ("max_reset_tries", C.c_int32)]


class RolloutArgs(C.Structure):
    _fields_ = [("n", C.c_int64), ("env_id0", C.c_int64), ("K", C.c_int32), ("n_grids", C.c_int32),
                ("envs_per_grid", C.c_int64), ("sq", C.c_void_p), ("sq_offsets", C.c_void_p),
                ("seed", C.c_uint64), ("actions_f32", C.c_void_p), ("reset_u", C.c_void_p),
                ("R", C.c_int32),
                ("j1", C.c_void_p), ("j2", C.c_void_p), ("reward", C.c_void_p), ("flags", C.c_void_p),
                ("step_ctr", C.c_void_p), ("reset_ctr", C.c_void_p), ("ep_len", C.c_void_p),
                ("rec_j1", C.c_void_p), ("rec_j2", C.c_void_p), ("rec_reward", C.c_void_p),
                ("rec_flags", C.c_void_p), ("stats", C.c_void_p), ("threads", C.c_int32)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ag_oracle.c")
    hdr = os.path.join(_HERE, "ag_oracle.h")
    if (force or not os.path.exists(_SO)
            or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["make", "-s", "-C", _HERE], env={k: v for k, v in os.environ.items()
                                                                if k not in ("CC", "CFLAGS")})
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.ago_grid_squares.restype = C.c_int64
        _lib.ago_experiment_loop.restype = C.c_int64
        _lib.ago_segment_square_margin.restype = C.c_double
        _lib.ago_rollout.restype = C.c_int
        _lib.ago_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_params() -> Params:
    p = Params()
    lib().ago_default_params(C.byref(p))
    return p


def num_threads() -> int:
    return int(lib().ago_num_threads())


def grid_squares(occ, env_size=1.6):
    """occupancy matrix -> (squares [M,4] f64 (min_x,min_y,max_x,max_y), cell_index [M] i32)."""
    occ8 = np.ascontiguousarray(np.asarray(occ) != 0, dtype=np.uint8)
    S = occ8.shape[0]
    assert occ8.shape == (S, S)
    m = int(occ8.sum())
    sq = np.zeros((max(m, 1), 4), dtype=np.float64)
    ci = np.zeros(max(m, 1), dtype=np.int32)
    got = lib().ago_grid_squares(_p(occ8), C.c_int32(S), C.c_double(env_size), _p(sq), _p(ci), C.c_int64(m))
    assert got == m
    return sq[:m], ci[:m]


def cells_to_squares(cols, rows, S, env_size=1.6):
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    sq = np.zeros((len(cols), 4), dtype=np.float64)
    lib().ago_cells_to_squares(_p(cols), _p(rows), C.c_int64(len(cols)), C.c_int32(S),
                               C.c_double(env_size), _p(sq))
    return sq


def manual_grid():
    """The hard-coded 3-obstacle 9x9 map, environment/occupancy_grid.py:45-50, in list order."""
    cols, rows = [6, 7, 3], [5, 5, 2]
    sq = cells_to_squares(cols, rows, 9)
    ci = np.array([r * 9 + c for c, r in zip(cols, rows)], dtype=np.int32)
    return sq, ci


def line_function(p0x, p0y, p1x, p1y):
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    lib().ago_line_function(C.c_double(p0x), C.c_double(p0y), C.c_double(p1x), C.c_double(p1y),
                            C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def segment_square(p0x, p0y, p1x, p1y, square, section_eps=1e-10):
    sq = np.ascontiguousarray(square, dtype=np.float64)
    aa = C.c_int64(0)
    r = lib().ago_segment_square(C.c_double(p0x), C.c_double(p0y), C.c_double(p1x), C.c_double(p1y),
                                 _p(sq), C.c_double(section_eps), C.byref(aa))
    return bool(r), aa.value


def forward_kinematics(j1, j2, l1=0.4, l2=0.3):
    o = [C.c_double() for _ in range(4)]
    lib().ago_forward_kinematics(C.c_double(j1), C.c_double(j2), C.c_double(l1), C.c_double(l2),
                                 *[C.byref(x) for x in o])
    return tuple(x.value for x in o)


def inverse_kinematics(tx, ty, l1=0.4, l2=0.3, corrected=False):
    """-> (valid, [j1_1, j2_1, j1_2, j2_2])"""
    sol = (C.c_double * 4)()
    lib().ago_inverse_kinematics.restype = C.c_int
    ok = lib().ago_inverse_kinematics(C.c_double(tx), C.c_double(ty), C.c_double(l1), C.c_double(l2),
                                      C.c_int(1 if corrected else 0), sol)
    return bool(ok), [sol[i] for i in range(4)]


def move_to_joint_pose(j1, j2, t1, t2, steps=100):
    a, b = C.c_double(j1), C.c_double(j2)
    lib().ago_move_to_joint_pose(C.byref(a), C.byref(b), C.c_double(t1), C.c_double(t2), C.c_int32(steps))
    return a.value, b.value


def collision_batch(j1, j2, squares, cell_index=None, params=None, want_first_hit=True, want_margin=False):
    p = params or default_params()
    j1 = np.ascontiguousarray(j1, dtype=np.float64)
    j2 = np.ascontiguousarray(j2, dtype=np.float64)
    sq = np.ascontiguousarray(squares, dtype=np.float64).reshape(-1, 4)
    ci = None if cell_index is None else np.ascontiguousarray(cell_index, dtype=np.int32)
    n = j1.shape[0]
    hit = np.zeros(n, dtype=np.uint8)
    fh = np.full(n, -1, dtype=np.int32) if want_first_hit else None
    mg = np.zeros(n, dtype=np.float64) if want_margin else None
    lib().ago_collision_batch(C.byref(p), _p(sq), _p(ci), C.c_int64(sq.shape[0]), _p(j1), _p(j2),
                              _p(hit), _p(fh), _p(mg), C.c_int64(n))
    return hit, fh, mg


def step_batch(j1, j2, actions, reward, flags, squares, cell_index=None, params=None,
               want_margin=False):
    """In-place batched Scene.step.  Returns dict(ee, dist, first_hit, margin, axis_aligned)."""
    p = params or default_params()
    for arr, dt in ((j1, np.float64), (j2, np.float64), (reward, np.float64), (flags, np.uint8)):
        assert arr.dtype == dt and arr.flags.c_contiguous
    actions = np.ascontiguousarray(actions, dtype=np.float64)
    sq = np.ascontiguousarray(squares, dtype=np.float64).reshape(-1, 4)
    ci = None if cell_index is None else np.ascontiguousarray(cell_index, dtype=np.int32)
    n = j1.shape[0]
    ee = np.zeros((n, 2)); dist = np.zeros((n, 2))
    fh = np.full(n, -1, dtype=np.int32)
    mg = np.zeros(n) if want_margin else None
    aa = C.c_int64(0)
    lib().ago_step_batch(C.byref(p), _p(sq), _p(ci), C.c_int64(sq.shape[0]), _p(j1), _p(j2), _p(actions),
                         _p(reward), _p(flags), _p(ee), _p(dist), _p(fh), _p(mg), C.byref(aa), C.c_int64(n))
    return dict(ee=ee, dist=dist, first_hit=fh, margin=mg, axis_aligned=aa.value)


def experiment_loop(j1, j2, draws, steps, squares, params=None):
    """experiment_0.thread_function for one env.  Returns (rec [steps,7], reset_steps, (j1,j2), draws_used)."""
    p = params or default_params()
    sq = np.ascontiguousarray(squares, dtype=np.float64).reshape(-1, 4)
    draws = np.ascontiguousarray(draws, dtype=np.float64)
    rec = np.zeros((steps, 7), dtype=np.float64)
    cj1, cj2 = C.c_double(j1), C.c_double(j2)
    nres = C.c_int64(0)
    rs = np.zeros(steps, dtype=np.int64)
    used = lib().ago_experiment_loop(C.byref(p), _p(sq), C.c_int64(sq.shape[0]), C.byref(cj1), C.byref(cj2),
                                     _p(draws), C.c_int64(draws.shape[0]), C.c_int64(steps), _p(rec),
                                     C.byref(nres), _p(rs), C.c_int64(steps))
    if used < 0:
        raise RuntimeError("draw stream exhausted")
    return rec, rs[:nres.value].copy(), (cj1.value, cj2.value), int(used)


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().ago_philox4x32_10(c, k, o)
    return tuple(int(x) for x in o)


def philox_uniform2(seed, env_id, draw, stream):
    u = (C.c_double * 2)()
    lib().ago_philox_uniform2(C.c_uint64(seed), C.c_uint64(env_id), C.c_uint32(draw), C.c_uint32(stream), u)
    return u[0], u[1]


class RolloutState:
    """Host mirror of the product's per-env SoA state."""

    def __init__(self, j1, j2):
        n = len(j1)
        self.j1 = np.ascontiguousarray(j1, dtype=np.float64).copy()
        self.j2 = np.ascontiguousarray(j2, dtype=np.float64).copy()
        self.reward = np.zeros(n, dtype=np.float32)
        self.flags = np.zeros(n, dtype=np.uint8)
        self.step_ctr = np.zeros(n, dtype=np.uint32)
        self.reset_ctr = np.zeros(n, dtype=np.uint32)
        self.ep_len = np.zeros(n, dtype=np.uint32)


def rollout(state, K, grids, envs_per_grid=None, env_id0=0, seed=0, actions_f32=None, reset_u=None,
            record=True, params=None, threads=0):
    """grids: list of squares arrays [M_g,4].  Returns (rec dict or None, stats int64[ST_COUNT])."""
    p = params or default_params()
    n = state.j1.shape[0]
    if not isinstance(grids, (list, tuple)):
        grids = [grids]
    offs = np.zeros(len(grids) + 1, dtype=np.int64)
    for i, g in enumerate(grids):
        offs[i + 1] = offs[i] + np.asarray(g).reshape(-1, 4).shape[0]
    allsq = np.ascontiguousarray(np.concatenate([np.asarray(g, dtype=np.float64).reshape(-1, 4) for g in grids])
                                 if offs[-1] else np.zeros((1, 4)))
    a = RolloutArgs()
    a.n, a.env_id0, a.K, a.n_grids = n, env_id0, K, len(grids)
    a.envs_per_grid = envs_per_grid or max(n, 1)
    a.sq, a.sq_offsets, a.seed = _p(allsq), _p(offs), seed
    if actions_f32 is not None:
        actions_f32 = np.ascontiguousarray(actions_f32, dtype=np.float32)
        assert actions_f32.shape == (K, n, 2)
    a.actions_f32 = _p(actions_f32)
    if reset_u is not None:
        reset_u = np.ascontiguousarray(reset_u, dtype=np.float64)
        assert reset_u.shape[0] == n and reset_u.shape[2] == 2
        a.R = reset_u.shape[1]
    a.reset_u = _p(reset_u)
    a.j1, a.j2, a.reward, a.flags = _p(state.j1), _p(state.j2), _p(state.reward), _p(state.flags)
    a.step_ctr, a.reset_ctr, a.ep_len = _p(state.step_ctr), _p(state.reset_ctr), _p(state.ep_len)
    rec = None
    if record:
        rec = dict(j1=np.zeros((K, n), np.float32), j2=np.zeros((K, n), np.float32),
                   reward=np.zeros((K, n), np.float32), flags=np.zeros((K, n), np.uint8))
        a.rec_j1, a.rec_j2 = _p(rec["j1"]), _p(rec["j2"])
        a.rec_reward, a.rec_flags = _p(rec["reward"]), _p(rec["flags"])
    stats = np.zeros(ST_COUNT, dtype=np.int64)
    a.stats = _p(stats)
    a.threads = threads
    rc = lib().ago_rollout(C.byref(p), C.byref(a))
    if rc != 0:
        raise RuntimeError("ago_rollout failed: %d" % rc)
    return rec, stats
