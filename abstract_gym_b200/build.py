"""In-tree build of the CUDA extension (libabstract_gym_b200.so) for sm_100a.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the gpurun snapshot.
Usage:  python -m abstract_gym_b200.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(HERE, "..", "include")
LIB = os.environ.get("AG_LIB_PATH") or os.path.join(HERE, "libabstract_gym_b200.so")   # AG_LIB_PATH: A/B builds
SOURCES = ["ag_kernels.cu", "ag_host.cu"]
HEADERS = ["ag_device.cuh", "ag_fast.cuh", os.path.join(INCLUDE, "abstract_gym_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
    "-Xptxas", "-v",
    "--shared",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if os.environ.get("AG_LIB_PATH"):
        return False
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str = None) -> str:
    if out is None and not force and not needs_build():
        return LIB
    extra = os.environ.get("AG_NVCC_EXTRA", "").split()          # A/B builds, e.g. -DAG_ROLLOUT_SMALL_BLOCK=64
    cmd = [find_nvcc()] + NVCC_FLAGS + extra + ["-I", INCLUDE, "-o", out or LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = (out + ".log") if out else os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed (%d); see %s" % (r.returncode, log))
    return out or LIB


if __name__ == "__main__":
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, out=out))
