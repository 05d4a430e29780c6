"""Builds csrc/libcausalstump_b200.so in-tree with nvcc for sm_100a (no CPU variant)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "libcausalstump_b200.so")
SOURCES = ["api.cu"]
# every source of the library: a header missing here means edits to it are silently not rebuilt
DEPS = sorted(f for f in os.listdir(HERE) if f.endswith((".cu", ".cuh", ".h"))) + [os.path.join("..", "..", "include", "cs_b200.h")]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libcausalstump_b200.so can only be built with the CUDA toolkit")


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(os.path.join(HERE, d)) <= t for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "--cudart", "static"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(HERE, s) for s in SOURCES] + ["-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libcausalstump_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
