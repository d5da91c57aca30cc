"""abstract_gym_b200 -- B200-native batched simulator for abstract_gym's scene_0 step/reset path.

The reference's class API is kept (same module layout under this package):
    utils.geometry.{Point, Line, Square}, utils.collision_checker.CollisionChecker,
    robot.two_joint_robot.TwoJointRobot, environment.occupancy_grid.OccupancyGrid,
    scenario.scene_0.Scene
with batched siblings (BatchedTwoJointRobot, BatchedOccupancyGrid, BatchedScene) whose state is
structure-of-arrays CUDA tensors.  All arithmetic of the path runs in hand-written sm_100a CUDA
kernels behind the C ABI of include/abstract_gym_b200.h; there is no CPU fallback.
"""
from . import _lib
from ._lib import AgError, ENGINES, STAT_NAMES, default_params, launch_count
from .utils.geometry import Point, Line, Square
from .utils.collision_checker import CollisionChecker, segment_square_arrays
from .robot.two_joint_robot import TwoJointRobot, BatchedTwoJointRobot, forward_kinematics
from .environment.occupancy_grid import OccupancyGrid, BatchedOccupancyGrid, DeviceGrid
from .scenario.scene_0 import Scene, BatchedScene
from .scenario.vector_env import VectorEnv
from . import ops          # registers torch.ops.abstract_gym_b200.{collision_check, step, reset, rollout}
from .experiment.experiment_0 import Trajectories, run_experiment, run_experiment_exact

__all__ = ["AgError", "ENGINES", "STAT_NAMES", "default_params", "launch_count", "Point", "Line", "Square",
           "CollisionChecker", "segment_square_arrays", "TwoJointRobot", "BatchedTwoJointRobot",
           "forward_kinematics", "OccupancyGrid", "BatchedOccupancyGrid", "DeviceGrid", "Scene", "BatchedScene", "VectorEnv", "Trajectories", "run_experiment",
           "run_experiment_exact"]
