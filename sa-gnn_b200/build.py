"""Builds lib/libsagnn_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SOURCES = ["csrc/plan.cu", "csrc/spmm.cu", "csrc/pairs.cu", "csrc/events.cu", "csrc/sampler.cu", "csrc/np_stream.cu"]
HEADERS = ["csrc/common.cuh", "csrc/spmm_rpw.cuh", "csrc/spmm_pkt.cuh", "../include/sagnn_b200.h"]
LIB = os.path.join(HERE, "lib", "libsagnn_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--threads", "4",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xcompiler", "-Wno-deprecated-declarations",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out: experiment variants, e.g. build(defines=["SAGNN_MIN_BLOCKS=4"], out="lib/x.so")."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(HERE, "csrc")]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-D" + d for d in defines]
    target = LIB if out is None else os.path.join(HERE, out)
    cmd += [os.path.join(HERE, s) for s in SOURCES] + ["-o", target]
    env = dict(os.environ)
    env.pop("CC", None)   # the image's CC points at a gcc wrapper without its spec files
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
