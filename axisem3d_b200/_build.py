"""Build the CUDA library in-tree (axisem3d_b200/libaxisem3d_b200.so) with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libaxisem3d_b200.so")
SRC = os.path.join(HERE, "csrc", "api.cu")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("api.cu", "kernels.cuh", "elem.cuh", "fft.cuh", "fused.cuh")] + [
    os.path.join(ROOT, "include", "axisem3d_b200.h")]


def nvcc_path():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False, with_nccl=True):
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), "-shared", "-Xcompiler", "-fPIC", "-O3", "-std=c++17", "-lineinfo",
           "-gencode", "arch=compute_100a,code=sm_100a", "-ftz=true", "-Xptxas", "-v" if verbose else "-O3",
           "-o", LIB, SRC]
    if with_nccl and os.path.exists("/usr/include/nccl.h"):
        cmd += ["-DAX3D_WITH_NCCL", "-lnccl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
