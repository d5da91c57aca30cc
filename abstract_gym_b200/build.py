"""In-tree build of the CUDA extension (libabstract_gym_b200.so) for sm_100a.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the gpurun snapshot.
Every translation unit is compiled to an object file in parallel, then linked.
Usage:  python -m abstract_gym_b200.build [--force] [--verbose] [--out PATH]
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(HERE, "..", "include")
LIB = os.environ.get("AG_LIB_PATH") or os.path.join(HERE, "libabstract_gym_b200.so")   # AG_LIB_PATH: A/B builds
SOURCES = ["ag_kernels.cu", "ag_rollout_lut.cu", "ag_dense.cu", "ag_host.cu"]
HEADERS = ["ag_device.cuh", "ag_fast.cuh", "ag_rollout.cuh", os.path.join(INCLUDE, "abstract_gym_b200.h")]
OBJDIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
    "-Xptxas", "-v",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build() -> bool:
    if os.environ.get("AG_LIB_PATH"):
        return False
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in sources()] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, out: str = None) -> str:
    if out is None and not force and not needs_build():
        return LIB
    nvcc = find_nvcc()
    extra = os.environ.get("AG_NVCC_EXTRA", "").split()          # A/B builds, e.g. -DAG_LUT_BITS=10
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    target = out or LIB
    objdir = OBJDIR if out is None else target + ".obj"
    os.makedirs(objdir, exist_ok=True)
    logs = {}

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-I", INCLUDE, "-c", "-o", obj, os.path.join(CSRC, src)]
        r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        logs[src] = " ".join(cmd) + "\n" + r.stdout
        return r.returncode, obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(sources())) as ex:
        results = list(ex.map(compile_one, sources()))
    rc = max(r for r, _ in results)
    link_out = ""
    if rc == 0:
        cmd = [nvcc, "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", target] + [o for _, o in results]
        r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        rc, link_out = r.returncode, " ".join(cmd) + "\n" + r.stdout
    log = (out + ".log") if out else os.path.join(HERE, "build.log")
    text = "".join(logs[s] for s in sources()) + link_out
    with open(log, "w") as f:
        f.write(text)
    if verbose or rc != 0:
        sys.stderr.write(text)
    if rc != 0:
        raise RuntimeError("nvcc failed (%d); see %s" % (rc, log))
    return target


if __name__ == "__main__":
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, out=out))
