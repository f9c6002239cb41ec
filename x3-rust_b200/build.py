"""Build libx3b200.so (hand-written sm_100a CUDA + the C ABI of include/x3_b200.h) in-tree with nvcc.

    python x3-rust_b200/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libx3b200.so")
SOURCES = ["x3_api.cu", "x3_encode.cu", "x3_decode.cu", "x3_synth.cu"]


def _newest_source_mtime():
    m = os.path.getmtime(os.path.join(HERE, "..", "include", "x3_b200.h"))
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh", ".h")):
            m = max(m, os.path.getmtime(os.path.join(CSRC, f)))
    return m


def build(force=False, verbose=False):
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= _newest_source_mtime():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", SO] + [os.path.join(CSRC, s) for s in SOURCES]
    cmd += os.environ.get("X3_NVCC_FLAGS", "").split()
    if verbose:
        cmd += ["-Xptxas", "-v"]
    subprocess.check_call(cmd)
    build_cli()
    return SO


def build_cli():
    """x3-rust_b200/host/x3 : the reference's CLI (src/bin/x3.rs) over the C++ host mirror (host/x3.hpp)."""
    host = os.path.join(HERE, "host")
    exe = os.path.join(host, "x3")
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-o", exe, os.path.join(host, "x3_cli.cpp"),
           "-L" + HERE, "-lx3b200", "-Wl,-rpath,$ORIGIN/.."]
    subprocess.check_call(cmd)
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
