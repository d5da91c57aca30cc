"""Vendor the UNMODIFIED reference into baseline/_ref/abstract_gym/ (git-ignored; travels to the GPU box with gpurun).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference is 787 lines of Python that import each other through
``sys.path`` hacks which need a checkout literally named ``abstract_gym``; /root/reference does
not exist on the GPU box, so ``__graft_entry__.build()`` copies it here (when present) and bench.py's cpu_baseline leg
times its loop (experiment/experiment_0.py:20-34) on the bench box's host cores next to the C port.  Nothing under
abstract_gym_b200/ reads this copy; the sources never enter the git history.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "..", "baseline", "_ref", "abstract_gym")
SRC = os.environ.get("ABSTRACT_GYM_REFERENCE_SRC", "/root/reference")


def vendored_root():
    """path of the vendored copy, or None"""
    p = os.path.abspath(DEST)
    return p if os.path.isfile(os.path.join(p, "scenario", "scene_0.py")) else None


def vendor(force: bool = False):
    """copy SRC -> baseline/_ref/abstract_gym if SRC exists; returns the vendored path or None"""
    if vendored_root() and not force:
        return vendored_root()
    if not os.path.isfile(os.path.join(SRC, "scenario", "scene_0.py")):
        return vendored_root()
    dest = os.path.abspath(DEST)
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    os.makedirs(os.path.dirname(dest), exist_ok=True)
    shutil.copytree(SRC, dest, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc"))
    return vendored_root()


if __name__ == "__main__":
    print(vendor(force=True))
