"""Gym-style batched front end of scene_0 for RL training loops: the callers of
`Scene.step` / `Scene.reset` (scenario/scene_0.py:88-113) want observations, rewards and termination flags
for N environments per call, with terminated environments restarted automatically the way
experiment/experiment_0.py:30-34 does.

    env = VectorEnv(4096, device="cuda")
    obs = env.reset()
    obs, reward, terminated, truncated, info = env.step(action)      # action: [N, 2] joint deltas

Observation [N, 6] float64: (joint_1, joint_2, EE_x, EE_y, |target_x - EE_x|, |target_y - EE_y|; per-env targets optional) -- the two
quantities `check_target_reached` thresholds (scene_0.py:129-130) are the reference's only notion of goal distance.
Every step is ONE launch of the fused kernel K6 (`ag_step_obs`: step, terminal observation, episode statistics,
Scene.reset() of terminated envs, next observation) on preallocated buffers, so `capture()` can record it once into a
CUDA graph and `step` replays it.
"""
import torch

from .. import _lib
from .._device import ptr, require_cuda, stream_ptr
from ..environment.occupancy_grid import OccupancyGrid
from ..robot.two_joint_robot import BatchedTwoJointRobot
from .scene_0 import BatchedScene


class VectorEnv:
    def __init__(self, num_envs, env=None, device=None, seed=0, engine="fast", auto_reset=True, target_c=None,
                 choose_j_tar=False, env_id0=0, targets=None, crop_size=0):
        dev = require_cuda(device)
        self.num_envs = int(num_envs)
        self.device = dev
        self.auto_reset = bool(auto_reset)
        grid = env if env is not None else OccupancyGrid(size=9, random_obstacle=False)      # scene_0's map
        gen = torch.Generator(device=dev).manual_seed(int(seed))
        robot = BatchedTwoJointRobot.random(self.num_envs, device=dev, generator=gen)
        self.scene = BatchedScene(robot, grid, target_c=target_c, engine=engine, seed=seed, env_id0=env_id0)
        self.scene.choose_j_tar = bool(choose_j_tar)
        n = self.num_envs
        self._action = torch.zeros(n, 2, dtype=torch.float64, device=dev)
        self._fk = torch.zeros(n, 4, dtype=torch.float64, device=dev)
        self._obs = torch.zeros(n, 6, dtype=torch.float64, device=dev)
        self._final_obs = torch.zeros(n, 6, dtype=torch.float64, device=dev)
        self._reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self._term_u8 = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._coll_u8 = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._terminated = self._term_u8.view(torch.bool)          # same bytes: the kernel writes 0 / 1
        self._collision = self._coll_u8.view(torch.bool)
        # optional local occupancy crop around the end effector: [N, c, c] uint8 (1 occupied, 2 outside the grid)
        self._crop = torch.zeros(n, int(crop_size), int(crop_size), dtype=torch.uint8, device=dev) if crop_size else None
        # per-env cartesian targets [N,2] (None: every env uses target_c); set_targets() may change them between steps
        self._targets = None
        if targets is not None:
            self._targets = torch.zeros(n, 2, dtype=torch.float64, device=dev)
            self.set_targets(targets)
        self._graph = None
        self._lib = _lib.load()

    def set_targets(self, targets):
        """per-env cartesian targets [N,2]; written into the buffer the (possibly captured) step reads"""
        if self._targets is None:
            raise RuntimeError("construct VectorEnv(targets=...) to use per-env targets")
        self._targets.copy_(torch.as_tensor(targets, dtype=torch.float64, device=self.device).reshape(self.num_envs, 2))

    # ---- one launch per step (K6, csrc/ag_kernels.cu k_step_obs) on preallocated buffers: capturable -----------
    def _step_impl(self):
        sc = self.scene
        g = sc.grid.c_struct()
        _lib.check(self._lib.ag_step_obs(sc.params(), g, ptr(sc.robot.joint_1), ptr(sc.robot.joint_2), ptr(self._action), 0,
                                         ptr(sc.step_reward), ptr(sc.flags), ptr(sc.reset_ctr), ptr(sc.ep_len),
                                         ptr(self._targets), ptr(self._obs), ptr(self._reward), ptr(self._term_u8),
                                         ptr(self._coll_u8), ptr(self._final_obs), ptr(self._crop),
                                         0 if self._crop is None else self._crop.shape[1], ptr(sc.stats), sc.seed,
                                         1 if self.auto_reset else 0, self.num_envs, sc.env_id0, sc.engine,
                                         stream_ptr(self.device)), "ag_step_obs")

    def _observe(self, out):
        """observation of the current poses (after reset()): FK over arrays, then the two goal distances"""
        sc = self.scene
        _lib.check(self._lib.ag_forward_kinematics(sc.params(), ptr(sc.robot.joint_1), ptr(sc.robot.joint_2),
                                                   ptr(self._fk), self.num_envs, stream_ptr(self.device)),
                   "ag_forward_kinematics")
        out[:, 0].copy_(sc.robot.joint_1); out[:, 1].copy_(sc.robot.joint_2)
        out[:, 2:4].copy_(self._fk[:, 2:4])
        if self._targets is None:
            out[:, 4].copy_((float(sc.target_c.x) - self._fk[:, 2]).abs())
            out[:, 5].copy_((float(sc.target_c.y) - self._fk[:, 3]).abs())
        else:
            out[:, 4:6].copy_((self._targets - self._fk[:, 2:4]).abs())

    # ---- public API ------------------------------------------------------------------------------
    def reset(self):
        """Scene.random_valid_pose + Scene.reset for every env (experiment_0.py:16); returns obs [N, 6]."""
        self.scene.random_valid_pose()
        self.scene.reset()
        self._observe(self._obs)
        return self._obs

    def capture(self):
        """Record one step (the single K6 launch) into a CUDA graph.  The scene constants (target_c, choose_j_tar, link
        lengths, engine) are frozen into the graph: change them and capture() again; per-env targets and the action
        buffer stay live."""
        torch.cuda.synchronize(self.device)
        snap = self.scene.state_dict()
        outs = [t.clone() for t in (self._obs, self._final_obs, self._reward, self._term_u8, self._coll_u8)]
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self._step_impl()                # warm-up outside capture
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._step_impl()
        self.scene.load_state_dict(snap)     # the warm-up and the capture advanced nothing the caller should see
        for t, o in zip((self._obs, self._final_obs, self._reward, self._term_u8, self._coll_u8), outs):
            t.copy_(o)                       # ... including the output buffers a caller may hold (reset() returns _obs)
        torch.cuda.synchronize(self.device)
        self._graph = graph
        return self

    def step(self, action):
        """action: [N, 2] joint deltas (any float dtype).  Returns (obs, reward, terminated, truncated, info);
        terminated = done | collision (experiment_0.py:30); truncated is always False (the reference has no time limit);
        info: collision [N] bool, final_obs [N, 6] (the terminal observation of envs that were restarted), crop (if
        requested) [N, c, c] uint8 occupancy around the end effector."""
        self._action.copy_(torch.as_tensor(action, device=self.device).reshape(self.num_envs, 2))
        if self._graph is not None:
            self._graph.replay()
        else:
            self._step_impl()
        return (self._obs, self._reward, self._terminated, torch.zeros_like(self._terminated),
                dict(collision=self._collision, final_obs=self._final_obs, crop=self._crop))

    def sample_action(self, scale_factor=0.1, generator=None):
        return self.scene.sample_action(scale_factor, generator)

    def stats(self):
        return self.scene.stats_dict()
