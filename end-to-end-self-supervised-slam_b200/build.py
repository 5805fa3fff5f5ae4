"""Build the C-ABI CUDA library for sm_100a, in-tree (so it travels to the GPU box with the snapshot).

    python end-to-end-self-supervised-slam_b200/build.py [--force] [--verbose]

Output: end-to-end-self-supervised-slam_b200/lib/libe2e_slam_b200.so  (git-ignored, NOT gpurun-ignored)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libe2e_slam_b200.so")
INCLUDE = os.path.join(HERE, "..", "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only; no PTX for other targets, no fallbacks
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
    # no --use_fast_math: the forward path is bit-exact with the reference (see csrc/common.cuh)
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "e2e_slam_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("E2E_NVCC_FLAGS", "").split()      # kernel-tuning experiments only
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-I", INCLUDE] + sources() + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libe2e_slam_b200.so")
    with open(os.path.join(LIBDIR, "ptxas.log"), "w") as f:
        f.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
