"""Build libmcq.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

One object per kernel family, compiled in parallel, linked into one shared library."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libmcq.so")
PTXAS_LOG = os.path.join(PKG, "libmcq.ptxas.log")
OBJ_DIR = os.path.join(HERE, "build")
SOURCES = ["mcq_api.cu", "launch_fast.cu", "launch_spec.cu", "launch_anneal.cu", "launch_wide.cu"]
HEADERS = ["anneal.cuh", "accept.cuh", "geometry.cuh", "spec.cuh", "fast.cuh", "wide.cuh", "philox.cuh", "launch.h",
           os.path.join("..", "..", "include", "mcq.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-diag-suppress", "550,177",
]


def needs_build(out=OUT):
    if not os.path.isfile(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS + ["build.py"])


def build(force=False, verbose=False, extra_flags=(), out=OUT, log=PTXAS_LOG):
    if not force and not needs_build(out):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tag = os.path.splitext(os.path.basename(out))[0]
    os.makedirs(OBJ_DIR, exist_ok=True)

    # ptxas statistics are always collected: the log next to the library is what tests/test_host_logic.py reads to
    # make sure no kernel spills (a silent change of ptxas' register choice once cost the thread-per-chain kernel 25 %)
    def compile_one(src):
        obj = os.path.join(OBJ_DIR, f"{tag}.{os.path.splitext(src)[0]}.o")
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-Xptxas", "-v", "-c", "-o", obj, src]
        proc = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
        return src, obj, proc

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        done = list(ex.map(compile_one, SOURCES))
    for src, _obj, proc in done:
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
            raise RuntimeError(f"nvcc failed compiling {src}")
    link = subprocess.run([nvcc, "-shared", "-o", out] + [obj for _s, obj, _p in done], cwd=HERE, capture_output=True, text=True)
    if link.returncode != 0:
        sys.stderr.write(link.stdout + link.stderr)
        raise RuntimeError("nvcc failed linking libmcq.so")
    with open(log, "w") as f:
        for _src, _obj, proc in done:
            f.write(proc.stderr)
    if verbose:
        for _src, _obj, proc in done:
            sys.stderr.write(proc.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
