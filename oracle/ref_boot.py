"""Boot the UNMODIFIED reference (read-only at /root/reference) inside this container.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` (fixture generator) and by the
optional live differential tests in ``tests/`` (skipped when /root/reference is absent, e.g. on
the GPU box).  Nothing under ``abstract_gym_b200/`` may import this module.

The reference only imports when:
  1. it is reachable as a package literally named ``abstract_gym``;
  2. a directory holding an ``__init__.py`` is on ``sys.path`` (its modules do ``import __init__``);
  3. ``matplotlib`` resolves (scenario/scene_0.py:2-4 imports it unconditionally; only touched
     when ``visualize=True``).
We satisfy (1)/(2) with a temp directory holding a symlink ``abstract_gym -> /root/reference`` (import
mode tolerates the symlink; nothing is copied into the repo) and (3) with in-memory stub modules.
"""
import os
import sys
import tempfile
import types

def _root() -> str:
    return os.environ.get("ABSTRACT_GYM_REFERENCE", "/root/reference")


REFERENCE_ROOT = _root()


def available() -> bool:
    return os.path.isfile(os.path.join(_root(), "scenario", "scene_0.py"))


def _stub_matplotlib():
    if "matplotlib" in sys.modules:
        return
    try:
        import matplotlib  # noqa: F401
        return
    except Exception:
        pass
    mpl = types.ModuleType("matplotlib")
    pyplot = types.ModuleType("matplotlib.pyplot")
    patches = types.ModuleType("matplotlib.patches")
    path = types.ModuleType("matplotlib.path")

    class Path:  # only referenced by Scene.occ_to_patch (visualisation, out of scope)
        MOVETO, LINETO, CLOSEPOLY = 1, 2, 79

        def __init__(self, *a, **k):
            pass

    path.Path = Path
    mpl.pyplot, mpl.patches, mpl.path = pyplot, patches, path
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": pyplot,
                        "matplotlib.patches": patches, "matplotlib.path": path})


_booted = None


def boot():
    """Return a namespace with the reference classes (Point, Line, Square, CollisionChecker,
    TwoJointRobot, OccupancyGrid, Scene)."""
    global _booted
    if _booted is not None:
        return _booted
    if not available():
        raise RuntimeError("reference tree not found at %s" % _root())
    _stub_matplotlib()
    parent = tempfile.mkdtemp(prefix="ag_ref_")
    os.symlink(_root(), os.path.join(parent, "abstract_gym"))
    sys.path.insert(0, parent)
    sys.path.insert(0, os.path.join(parent, "abstract_gym"))  # makes ``import __init__`` resolve
    sys.dont_write_bytecode = True  # /root/reference is read-only
    import numpy as np
    st = np.random.get_state()  # importing the reference draws import-time randoms
    from abstract_gym.utils.geometry import Point, Line, Square
    from abstract_gym.utils.collision_checker import CollisionChecker
    from abstract_gym.robot.two_joint_robot import TwoJointRobot
    from abstract_gym.environment.occupancy_grid import OccupancyGrid
    from abstract_gym.scenario.scene_0 import Scene
    np.random.set_state(st)
    _booted = types.SimpleNamespace(Point=Point, Line=Line, Square=Square,
                                    CollisionChecker=CollisionChecker, TwoJointRobot=TwoJointRobot,
                                    OccupancyGrid=OccupancyGrid, Scene=Scene)
    return _booted
