"""torch custom ops over the C ABI (`torch.ops.abstract_gym_b200.*`).

Thin by design: each op checks device / dtype / contiguity, fetches the current CUDA stream and calls the
`extern "C"` symbol of include/abstract_gym_b200.h through ctypes.  There is no CPU dispatch: a CPU tensor is an
error.  The ops mutate their state tensors in place (declared via `mutates_args`) and return nothing, so they
can sit inside captured / compiled graphs next to a policy network.

    torch.ops.abstract_gym_b200.step(params, grid_bits, min_x, min_y, grid_meta, j1, j2, actions, reward, flags, stats, env_id0, engine)
    torch.ops.abstract_gym_b200.rollout(...)

`params` is the 11 float64 + 2 int32 `ag_params` struct packed as a float64[13] CPU tensor (see `pack_params`);
`grid_meta` = float64[8] CPU tensor (side, env_size, S, words_per_row, n_grids, max_occupied, grid_stride_words,
envs_per_grid).  The object API (`BatchedScene`) is the convenient front end; these ops are the functional one.
"""
import ctypes as C

import torch
from torch.library import custom_op

from . import _lib
from ._device import stream_ptr

NS = "abstract_gym_b200"


def pack_params(p: _lib.Params) -> torch.Tensor:
    return torch.tensor([p.link_1, p.link_2, p.target_x, p.target_y, p.target_j1, p.target_j2, p.reach_eps,
                         p.section_eps, p.reward_collision, p.reward_reach, p.action_scale, float(p.choose_j_tar),
                         float(p.max_reset_tries)], dtype=torch.float64)


def _params(t: torch.Tensor) -> _lib.Params:
    v = t.tolist()
    p = _lib.Params()
    (p.link_1, p.link_2, p.target_x, p.target_y, p.target_j1, p.target_j2, p.reach_eps, p.section_eps,
     p.reward_collision, p.reward_reach, p.action_scale) = v[:11]
    p.choose_j_tar, p.max_reset_tries = int(v[11]), int(v[12])
    return p


def pack_grid_meta(dg) -> torch.Tensor:
    """dg: abstract_gym_b200.DeviceGrid"""
    return torch.tensor([dg.side, dg.environment_size, dg.S, dg.words_per_row, dg.n_grids, dg.max_occupied,
                         dg.stride_words, dg.envs_per_grid], dtype=torch.float64)


def _grid(bits, min_x, min_y, meta) -> _lib.Grid:
    m = meta.tolist()
    g = _lib.Grid()
    g.bits, g.min_x, g.min_y = bits.data_ptr(), min_x.data_ptr(), min_y.data_ptr()
    g.side, g.env_size = m[0], m[1]
    g.S, g.words_per_row, g.n_grids, g.max_occupied = int(m[2]), int(m[3]), int(m[4]), int(m[5])
    g.grid_stride_words, g.envs_per_grid = int(m[6]), int(m[7])
    return g


def _dev(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if t.device.type != "cuda":
            raise RuntimeError("abstract_gym_b200 ops run on CUDA tensors only (no CPU dispatch)")
        if not t.is_contiguous():
            raise RuntimeError("abstract_gym_b200 ops need contiguous tensors")
        dev = t.device if dev is None else dev
        if t.device != dev:
            raise RuntimeError("all tensors must live on one device")
    return dev


def _need(t, dtype, name):
    if t.dtype != dtype:
        raise RuntimeError("%s must be %s, got %s" % (name, dtype, t.dtype))


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


@custom_op(NS + "::collision_check", mutates_args=("hit",))
def collision_check(params: torch.Tensor, grid_bits: torch.Tensor, min_x: torch.Tensor, min_y: torch.Tensor,
                    grid_meta: torch.Tensor, j1: torch.Tensor, j2: torch.Tensor, hit: torch.Tensor, env_id0: int,
                    engine: int) -> None:
    """Scene.collision_check (scenario/scene_0.py:60-76) for every env; hit: uint8 [N]"""
    dev = _dev(grid_bits, min_x, min_y, j1, j2, hit)
    _need(j1, torch.float64, "j1"); _need(j2, torch.float64, "j2"); _need(hit, torch.uint8, "hit")
    _lib.check(_lib.load().ag_collision_check(_params(params), _grid(grid_bits, min_x, min_y, grid_meta), _p(j1), _p(j2),
                                              _p(hit), None, j1.numel(), env_id0, engine, stream_ptr(dev)),
               "ag_collision_check")


@custom_op(NS + "::step", mutates_args=("j1", "j2", "reward", "flags", "stats"))
def step(params: torch.Tensor, grid_bits: torch.Tensor, min_x: torch.Tensor, min_y: torch.Tensor,
         grid_meta: torch.Tensor, j1: torch.Tensor, j2: torch.Tensor, actions: torch.Tensor, reward: torch.Tensor,
         flags: torch.Tensor, stats: torch.Tensor, env_id0: int, engine: int) -> None:
    """Scene.step (scenario/scene_0.py:88-103) for every env; actions [N,2] float64 or float32"""
    dev = _dev(grid_bits, min_x, min_y, j1, j2, actions, reward, flags, stats)
    _need(j1, torch.float64, "j1"); _need(j2, torch.float64, "j2"); _need(reward, torch.float32, "reward")
    _need(flags, torch.uint8, "flags"); _need(stats, torch.int64, "stats")
    if actions.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("actions must be float32 or float64")
    _lib.check(_lib.load().ag_step(_params(params), _grid(grid_bits, min_x, min_y, grid_meta), _p(j1), _p(j2), _p(actions),
                                   1 if actions.dtype == torch.float32 else 0, _p(reward), _p(flags), None, None, None,
                                   _p(stats), None, j1.numel(), env_id0, engine, stream_ptr(dev)), "ag_step")


@custom_op(NS + "::reset", mutates_args=("j1", "j2", "reward", "flags", "reset_ctr", "stats"))
def reset(params: torch.Tensor, grid_bits: torch.Tensor, min_x: torch.Tensor, min_y: torch.Tensor,
          grid_meta: torch.Tensor, j1: torch.Tensor, j2: torch.Tensor, reward: torch.Tensor, flags: torch.Tensor,
          reset_ctr: torch.Tensor, mask: torch.Tensor, stats: torch.Tensor, seed: int, env_id0: int,
          engine: int) -> None:
    """Scene.reset (scenario/scene_0.py:105-113) for envs with mask != 0, Philox reset candidates"""
    dev = _dev(grid_bits, min_x, min_y, j1, j2, reward, flags, reset_ctr, mask, stats)
    _need(mask, torch.uint8, "mask"); _need(reset_ctr, torch.int32, "reset_ctr")
    _lib.check(_lib.load().ag_reset(_params(params), _grid(grid_bits, min_x, min_y, grid_meta), _p(j1), _p(j2), _p(reward),
                                    _p(flags), _p(reset_ctr), _p(mask), None, 0, seed, 1, _p(stats), j1.numel(), env_id0,
                                    engine, stream_ptr(dev)), "ag_reset")


@custom_op(NS + "::rollout", mutates_args=("j1", "j2", "reward", "flags", "step_ctr", "reset_ctr", "ep_len", "rec_j1",
                                           "rec_j2", "rec_reward", "rec_flags", "stats"))
def rollout(params: torch.Tensor, grid_bits: torch.Tensor, min_x: torch.Tensor, min_y: torch.Tensor,
            grid_meta: torch.Tensor, j1: torch.Tensor, j2: torch.Tensor, reward: torch.Tensor, flags: torch.Tensor,
            step_ctr: torch.Tensor, reset_ctr: torch.Tensor, ep_len: torch.Tensor, actions: torch.Tensor,
            rec_j1: torch.Tensor, rec_j2: torch.Tensor, rec_reward: torch.Tensor, rec_flags: torch.Tensor,
            stats: torch.Tensor, seed: int, env_id0: int, engine: int) -> None:
    """The loop body of experiment/experiment_0.py:20-34 fused over K = actions.shape[0] steps (K4);
    actions [K,N,2] float32, rec_* [K,N]"""
    dev = _dev(grid_bits, min_x, min_y, j1, j2, reward, flags, step_ctr, reset_ctr, ep_len, actions, rec_j1, rec_j2,
               rec_reward, rec_flags, stats)
    _need(actions, torch.float32, "actions")
    n, K = j1.numel(), actions.shape[0]
    if tuple(actions.shape) != (K, n, 2) or tuple(rec_j1.shape) != (K, n):
        raise RuntimeError("actions must be [K,N,2] and records [K,N]")
    a = _lib.RolloutArgs()
    a.n, a.env_id0, a.K, a.engine, a.seed = n, env_id0, K, engine, seed
    a.actions, a.R = actions.data_ptr(), 0
    a.j1, a.j2, a.reward, a.flags = j1.data_ptr(), j2.data_ptr(), reward.data_ptr(), flags.data_ptr()
    a.step_ctr, a.reset_ctr, a.ep_len = step_ctr.data_ptr(), reset_ctr.data_ptr(), ep_len.data_ptr()
    a.rec_j1, a.rec_j2, a.rec_reward, a.rec_flags = (rec_j1.data_ptr(), rec_j2.data_ptr(), rec_reward.data_ptr(),
                                                     rec_flags.data_ptr())
    a.stats = stats.data_ptr()
    _lib.check(_lib.load().ag_rollout(_params(params), _grid(grid_bits, min_x, min_y, grid_meta), C.byref(a),
                                      stream_ptr(dev)), "ag_rollout")
