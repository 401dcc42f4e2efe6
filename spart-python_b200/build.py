#!/usr/bin/env python
"""Build libspart_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "csrc" / "spart_kernels.cu"
DEPS = sorted((HERE / "csrc").glob("*")) + [HERE.parent / "include" / "spart_b200.h"]
OUT = HERE / "spart_b200" / "lib" / "libspart_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def build(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if OUT.exists():          # GPU box without a toolkit on PATH: use the shipped binary
            return OUT
        raise RuntimeError("nvcc not found and no prebuilt libspart_b200.so")
    if OUT.exists() and not force and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in DEPS):
        return OUT
    OUT.parent.mkdir(parents=True, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(OUT), str(SRC)]
    print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
