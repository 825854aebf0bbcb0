"""Build ``libbfl.so`` (the C-ABI + CUDA kernels) in-tree for sm_100a with nvcc.

    python -m bayesflare_b200.build [--force]

nvcc cross-compiles without a GPU; the built ``.so`` is git-ignored but travels to the GPU box
with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libbfl.so")
SOURCES = ["bfl_kernels.cu", "bfl_wsplit.cu", "bfl_abi.cu"]
DEPS = SOURCES + ["bfl_kernels.cuh", "bfl_devmath.cuh", "phi_table.inc", "exphi_table.inc", "exphi2_table.inc", os.path.join(ROOT, "include", "bfl.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def cuda_lib_dir() -> str:
    return os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(nvcc_path()))), "lib64")


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    for d in DEPS:
        p = d if os.path.isabs(d) else os.path.join(CSRC, d)
        if os.path.getmtime(p) > t:
            return False
    return True


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    cmd = [nvcc_path(), "-std=c++17", "-O3", "-lineinfo",
           "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-shared",
           "-I", os.path.join(ROOT, "include"),
           "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lcufft", "-Xlinker", "-rpath=" + cuda_lib_dir()]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (res.stdout, res.stderr))
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
