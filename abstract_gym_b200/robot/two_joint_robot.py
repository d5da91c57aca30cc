"""Two-joint planar arm: the reference's `TwoJointRobot` (robot/two_joint_robot.py:7-113) and its
batched sibling whose state lives in HBM as structure-of-arrays float64 tensors.

Hot-path members (state, move_delta, end_effector, elbow_point) run through the CUDA extension.
`move_to_joint_pose`, `cart_target_valid_check` and `inverse_kinematic` are never called by
step/reset; they are kept as host code for interface completeness.
"""
import numpy as np
import torch

from .. import _lib
from .._device import as_f64, ptr, require_cuda, stream_ptr
from ..utils.geometry import Point


def _params_for(link_1, link_2):
    p = _lib.default_params()
    p.link_1, p.link_2 = float(link_1), float(link_2)
    return p


def forward_kinematics(j1, j2, link_1=0.4, link_2=0.3, device=None):
    """[n] joints -> [n,4] (elbow_x, elbow_y, ee_x, ee_y) float64 CUDA tensor
    (robot/two_joint_robot.py:31-47, unfused float64)."""
    dev = require_cuda(device if device is not None else (j1.device if torch.is_tensor(j1) else None))
    j1 = as_f64(j1, dev).reshape(-1)
    j2 = as_f64(j2, dev, j1.numel()).reshape(-1)
    out = torch.empty(j1.numel(), 4, dtype=torch.float64, device=dev)
    p = _params_for(link_1, link_2)
    _lib.check(_lib.load().ag_forward_kinematics(p, ptr(j1), ptr(j2), ptr(out), j1.numel(), stream_ptr(dev)),
               "ag_forward_kinematics")
    return out


class TwoJointRobot:
    """Single arm with the reference's constructor and methods.

    The reference evaluates its random default joints once at import time
    (robot/two_joint_robot.py:10-11), so every default-constructed robot shares one pose; here an
    omitted joint is drawn from np.random at construction time instead (same distribution)."""

    def __init__(self, joint_1=None, joint_2=None, link_1=0.4, link_2=0.3):
        self.joint_1 = np.random.rand() * np.pi * 2.0 if joint_1 is None else joint_1
        self.joint_2 = np.random.rand() * np.pi * 2.0 if joint_2 is None else joint_2
        self.link_1 = link_1
        self.link_2 = link_2
        self.EE = None        # stale caches in the reference too (refreshed by render only)
        self.elbow_p = None

    def total_length(self):
        return self.link_1 + self.link_2

    def _fk(self):
        return forward_kinematics([self.joint_1], [self.joint_2], self.link_1, self.link_2)[0].tolist()

    def end_effector(self):
        """robot/two_joint_robot.py:31-38"""
        _, _, x, y = self._fk()
        return Point(x, y)

    def elbow_point(self):
        """robot/two_joint_robot.py:40-47"""
        x, y, _, _ = self._fk()
        return Point(x, y)

    def move_delta(self, d1, d2):
        """robot/two_joint_robot.py:64-72 (no wrapping, no limits)"""
        self.joint_1 += d1
        self.joint_2 += d2

    # ---- not on the step/reset path -----------------------------------------------------------
    def move_to_joint_pose(self, target_j1, target_j2, steps=100):
        """robot/two_joint_robot.py:49-62: `steps` equal increments, no intermediate collision checks."""
        inc_1 = (1.0 / steps) * (target_j1 - self.joint_1)
        inc_2 = (1.0 / steps) * (target_j2 - self.joint_2)
        for _ in range(steps):
            self.joint_1 += inc_1
            self.joint_2 += inc_2

    def cart_target_valid_check(self, target_c):
        """robot/two_joint_robot.py:74-86: reachable annulus |l1-l2| < r <= l1+l2."""
        radius = np.sqrt(pow(target_c.x, 2) + pow(target_c.y, 2))
        return bool(abs(self.link_1 - self.link_2) < radius <= self.total_length()), radius

    def inverse_kinematic(self, target_c):
        """robot/two_joint_robot.py:88-113 (law of cosines).  Like the reference, alpha = arccos(x/r)
        drops the sign of target y; pass corrected=True to
        `inverse_kinematic_atan2` for the fixed variant."""
        return self._ik(target_c, corrected=False)

    def inverse_kinematic_atan2(self, target_c):
        return self._ik(target_c, corrected=True)

    def _ik(self, target_c, corrected):
        valid, radius = self.cart_target_valid_check(target_c)
        if not valid:
            print("Target out of reach.")
            return None, None
        if radius == 0:
            print("link_1 equals link 2 and the target is at origin, infinite many solutions.")
            return None, None
        l1, l2 = self.link_1, self.link_2
        theta = np.arccos((radius ** 2 + l1 ** 2 - l2 ** 2) / (2.0 * l1 * radius))
        alpha = np.arctan2(target_c.y, target_c.x) if corrected else np.arccos(target_c.x / radius)
        inner = np.pi - np.arccos((l1 ** 2 + l2 ** 2 - radius ** 2) / (2.0 * l1 * l2))
        first = np.array([alpha - theta, inner + (alpha - theta)])
        second = np.array([alpha + theta, (alpha + theta) - inner])
        return first, second


class BatchedTwoJointRobot:
    """N arms, SoA: joint_1[N], joint_2[N] float64 CUDA tensors (16 B per env)."""

    def __init__(self, joint_1, joint_2, link_1=0.4, link_2=0.3, device=None):
        dev = require_cuda(device if device is not None else (joint_1.device if torch.is_tensor(joint_1) else None))
        self.device = dev
        self.joint_1 = as_f64(joint_1, dev).reshape(-1).clone()
        self.joint_2 = as_f64(joint_2, dev, self.joint_1.numel()).reshape(-1).clone()
        self.link_1 = link_1
        self.link_2 = link_2

    @classmethod
    def random(cls, n, link_1=0.4, link_2=0.3, device=None, generator=None):
        """uniform poses j = (u*pi)*2.0, the map of scenario/scene_0.py:180-181"""
        dev = require_cuda(device)
        u = torch.rand(2, n, dtype=torch.float64, device=dev, generator=generator)
        j = (u * np.pi) * 2.0
        return cls(j[0], j[1], link_1, link_2, dev)

    def __len__(self):
        return self.joint_1.numel()

    def total_length(self):
        return self.link_1 + self.link_2

    def forward_kinematics(self):
        return forward_kinematics(self.joint_1, self.joint_2, self.link_1, self.link_2, self.device)

    def end_effector(self):
        """[N,2] float64"""
        return self.forward_kinematics()[:, 2:4]

    def elbow_point(self):
        """[N,2] float64"""
        return self.forward_kinematics()[:, 0:2]

    def move_delta(self, d1, d2):
        self.joint_1 += as_f64(d1, self.device)
        self.joint_2 += as_f64(d2, self.device)

    # ---- not on the step/reset path ---------------------------------------------------
    def move_to_joint_pose(self, target, steps=100):
        """robot/two_joint_robot.py:49-62 for every arm: target [N,2] joint poses, `steps` equal increments."""
        t = as_f64(target, self.device).reshape(len(self), 2).contiguous()
        _lib.check(_lib.load().ag_move_to_joint_pose(ptr(self.joint_1), ptr(self.joint_2), ptr(t), int(steps), len(self),
                                                     stream_ptr(self.device)), "ag_move_to_joint_pose")

    def inverse_kinematic(self, target_c, corrected=False):
        """robot/two_joint_robot.py:88-113 for M cartesian targets [M,2] -> (valid [M] bool, s1 [M,2], s2 [M,2]).
        Unreachable targets (the reference returns None, None) have valid == False and zero solutions.
        corrected=True uses atan2(y, x) for alpha; the reference's arccos(x/r) loses the sign of y."""
        t = as_f64(target_c, self.device).reshape(-1, 2).contiguous()
        m = t.shape[0]
        sol = torch.empty(m, 4, dtype=torch.float64, device=self.device)
        valid = torch.empty(m, dtype=torch.uint8, device=self.device)
        _lib.check(_lib.load().ag_inverse_kinematics(_params_for(self.link_1, self.link_2), ptr(t), ptr(sol), ptr(valid),
                                                     1 if corrected else 0, m, stream_ptr(self.device)),
                   "ag_inverse_kinematics")
        return valid != 0, sol[:, 0:2], sol[:, 2:4]
