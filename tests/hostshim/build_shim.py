"""Compile tests/hostshim/shim.cpp (the product's host+device maths headers, built for the CPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "libgms_hostshim.so")
SRC = os.path.join(HERE, "shim.cpp")
CSRC = os.path.join(HERE, "..", "..", "gaussian-mesh-splatting_b200", "csrc")


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("gms_common.cuh", "gms_preprocess.cuh", "gms_expand.cuh")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    fma = "-mfma" if " fma " in open("/proc/cpuinfo").read() else ""
    cmd = ["/usr/bin/g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", OUT, SRC]
    if fma:
        cmd.insert(2, fma)
    subprocess.check_call(cmd)
    return OUT
