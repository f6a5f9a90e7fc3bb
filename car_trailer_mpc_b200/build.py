"""In-tree build of the CUDA library ``libttmpc.so`` (sm_100a only).

``python -m car_trailer_mpc_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a
GPU; the resulting .so is git-ignored but travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shlex
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
SRC = os.path.join(PKG_DIR, "csrc", "ttmpc.cu")
DEPS = [SRC, os.path.join(PKG_DIR, "csrc", "ttmpc_core.cuh"), os.path.join(PKG_DIR, "csrc", "ttmpc_obca.cuh"),
        os.path.join(PKG_DIR, "csrc", "ttmpc_team.cuh"),
        os.path.join(ROOT, "include", "ttmpc.h")]
LIB = os.path.join(PKG_DIR, "libttmpc.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-diag-suppress", "128",
    "-shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def nvcc_path() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    out = os.environ.get("TTMPC_BUILD_OUT") or LIB  # experiment builds are judged by their own output file
    if not os.path.exists(out):
        return True
    if os.environ.get("TTMPC_NVCC_FLAGS") and out == LIB:
        return True  # extra flags aimed at the shipped library: always rebuild
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    # TTMPC_NVCC_FLAGS: extra flags for experiment builds (e.g. -DTTMPC_SPECULATION=1), TTMPC_BUILD_OUT: their output path
    out = os.environ.get("TTMPC_BUILD_OUT") or LIB
    cmd = [nvcc_path(), *NVCC_FLAGS, *shlex.split(os.environ.get("TTMPC_NVCC_FLAGS", "")), "-o", out, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
