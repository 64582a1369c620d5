"""Build libcrl_b200.so (the C-ABI library, include/crl_b200.h) in-tree with nvcc for sm_100a.

    python competitive-rl_b200/build.py [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcrl_b200.so")
SOURCES = ["crl_abi.cu", "pong_step.cu", "pong_raster.cu", "pong_raster_fast.cu", "crl_car_abi.cu",
           "car_physics.cu", "car_raster.cu"]
HEADERS = ["pong_common.cuh", "pong_raster_dev.cuh", "car_common.cuh", "car_contact.cuh", os.path.join("..", "..", "include", "crl_b200.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false",            # cv2's area resize and the fp64 ball physics are un-fused mul/add
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, f) for f in SOURCES] + ["-o", OUT]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
