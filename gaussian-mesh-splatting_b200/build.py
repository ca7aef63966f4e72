"""Build libgms_b200.so for sm_100a, in-tree (the .so travels to the GPU box with the repo snapshot).

    python gaussian-mesh-splatting_b200/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "gms_kernels.cu")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("gms_kernels.cu", "gms_common.cuh", "gms_preprocess.cuh",
                                                "gms_expand.cuh", "gms_composite_common.cuh", "gms_composite_fwd.cuh", "gms_composite_bwd.cuh", "gms_loss.cuh", "gms_sort.cuh", "gms_binning.cuh", "gms_image.cuh")] + \
       [os.path.join(HERE, "..", "include", "gms_b200.h")]
OUT = os.path.join(HERE, "gms_b200", "libgms_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
         "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(OUT)
