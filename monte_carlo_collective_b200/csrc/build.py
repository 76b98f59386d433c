"""Build libmcq.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libmcq.so")
PTXAS_LOG = os.path.join(PKG, "libmcq.ptxas.log")
SOURCES = ["mcq_api.cu"]
HEADERS = ["anneal.cuh", "spec.cuh", "wide.cuh", "philox.cuh", os.path.join("..", "..", "include", "mcq.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "550,177",
]


def needs_build():
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS + ["build.py"])


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # ptxas statistics are always collected: the log next to the library is what tests/test_host_logic.py reads to
    # make sure no kernel spills (a silent change of ptxas' register choice once cost the thread-per-chain kernel 25 %)
    cmd = [nvcc] + NVCC_FLAGS + ["-Xptxas", "-v", "-o", OUT] + SOURCES
    proc = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building libmcq.so")
    with open(PTXAS_LOG, "w") as f:
        f.write(proc.stderr)
    if verbose:
        sys.stderr.write(proc.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
