"""Throughput sweep over kernel geometry (development tool; writes gpurun_out/sweep.jsonl)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
import monte_carlo_collective_b200 as mcq  # noqa: E402
from monte_carlo_collective_b200 import schedules  # noqa: E402

eng = mcq.Engine(0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "sweep.jsonl"), "a")
LIN = {"type": "linear_annealing", "beta_start": 1.0, "beta_end": 3.0}


def run(tag, mode, n, nc, ns, history="none", **kw):
    import torch
    seeds = torch.arange(nc, dtype=torch.int64).cuda()
    best = None
    for _ in range(2):
        t0 = time.time()
        r = eng.run(mode, n, ns, seeds, schedules=LIN, history=history, device_buffers=True,
                    want_states=False, **kw)
        wall = time.time() - t0
        pps = nc * ns / (r.kernel_ms * 1e-3)
        if best is None or pps > best["pps"]:
            best = dict(tag=tag, mode=mode, n=n, chains=nc, steps=ns, history=history, kernel_ms=r.kernel_ms,
                        wall_ms=wall * 1e3, pps=pps, acc=float(r.n_accepted.float().mean()) / ns, **kw)
    print(json.dumps(best), flush=True)
    out.write(json.dumps(best) + "\n")
    out.flush()


which = sys.argv[1] if len(sys.argv) > 1 else "base"
if which == "quick32":
    run("q", "full_3d", 12, 20480, 200000, algo="table")
    run("q", "board", 12, 20480, 100000, algo="table")
elif which == "base":
    for mode in ("full_3d", "board"):
        for G in (4, 8, 16, 32):
            run("G", mode, 12, 20480, 20000, lanes_per_chain=G)
    for hist in ("full", "stats"):
        run("hist", "full_3d", 12, 20480, 20000, history=hist, lanes_per_chain=8)
    for mode in ("full_3d", "board"):
        for n in (8, 15, 20):
            run("N", mode, n, 8192, 10000)
    run("N", "board", 64, 1184, 2000)
elif which == "c2geo":
    run("c2", "full_3d", 12, 20480, 200000, algo="table")
    for lanes, m in ((32, 32), (32, 28), (16, 64), (16, 48), (16, 56)):
        run("c2", "full_3d", 12, 20480, 200000, algo="table", lanes_per_chain=lanes, max_chains_per_sm=m)
elif which == "auto":
    run("auto", "full_3d", 12, 20480, 200000, algo="table")
    run("auto", "board", 12, 20480, 100000, algo="table")
    run("auto", "full_3d", 12, 4096, 100000, algo="table")
    run("auto", "full_3d", 12, 1000, 100000, algo="table")
    run("auto", "full_3d", 12, 65536, 50000, algo="table")
    run("auto", "full_3d", 8, 20480, 100000, algo="table")
    run("auto", "board", 15, 8192, 50000, algo="table")
    run("auto", "board", 20, 8192, 20000, algo="table")
elif which == "quick":
    for lanes in (32, 16):
        run("q", "full_3d", 12, 20480, 200000, algo="table", lanes_per_chain=lanes)
        run("q", "board", 12, 20480, 100000, algo="table", lanes_per_chain=lanes)
elif which == "occ2":
    for lanes in (32, 16):
        cpw = 32 // lanes
        for ctas in (4, 5, 6, 7, 8):
            chains = 148 * ctas * 4 * cpw
            run("occ2", "full_3d", 12, chains, 100000, algo="table", lanes_per_chain=lanes, max_chains_per_sm=ctas * 4 * cpw)
elif which == "table":
    for mode in ("full_3d", "board"):
        for w in (1, 2, 4):
            run("table", mode, 12, 20480, 20000, algo="table", warps_per_cta=w)
    run("table", "full_3d", 12, 20480, 200000, algo="table")
    run("table", "full_3d", 12, 20480, 20000, algo="table", history="stats")
    for mode in ("full_3d", "board"):
        for n in (8, 15, 19):
            run("tableN", mode, n, 8192, 10000, algo="table")
    run("tableN", "board", 20, 8192, 10000, algo="table")
elif which == "occ":
    for G in (4, 8, 16):
        for m in (8, 16, 24, 32, 40, 48):
            run("occ", "full_3d", 12, 148 * m, 20000, lanes_per_chain=G, max_chains_per_sm=m)

if which == "tail":
    for st in ("1", "2"):
        os.environ["MCQ_STREAMS"] = st
        for nc in (4736 * 4, 20480, 4736 * 5, 4736 * 4 + 148 * 4):
            run("tail" + st, "full_3d", 12, nc, 100000, algo="table", lanes_per_chain=32)

if which == "bign":
    for mode, n, nc, ns in (("board", 64, 1184, 20000), ("board", 64, 16384, 20000), ("board", 64, 65536, 5000), ("board", 33, 8192, 20000),
                            ("board", 24, 8192, 20000), ("full_3d", 24, 8192, 20000), ("full_3d", 40, 4096, 10000)):
        for algo in ("lines", "gmem"):
            run("bign", mode, n, nc, ns, algo=algo)

if which == "gm":
    for mode, n, nc, ns in (("board", 64, 16384, 20000), ("board", 64, 65536, 5000), ("board", 33, 8192, 20000), ("full_3d", 40, 4096, 10000)):
        run("gm", mode, n, nc, ns, algo="gmem")

if which == "gmprof":
    run("gmprof", "board", 64, 16384, 3000, algo="gmem")

if which == "wide":
    for mode, n, nc, ns in (("board", 64, 1184, 20000), ("board", 64, 1184, 200000), ("board", 33, 1184, 100000), ("full_3d", 40, 1184, 50000)):
        run("wide", mode, n, nc, ns, algo="wide")
    run("wide-gmem", "board", 64, 16384, 20000, algo="gmem")

if which == "widex":
    for nc in (148, 1184, 4736, 16384):
        for algo in ("wide", "gmem"):
            run("widex", "board", 64, nc, 100000, algo=algo)
    run("widex", "board", 64, 148, 100000, algo="lines")
    for nc in (148, 1184, 8192):
        for algo in ("wide", "gmem", "lines"):
            run("widex", "board", 30, nc, 100000, algo=algo)

if which == "wideprof":
    run("wideprof", "board", 64, 148, 50000, algo="wide")

if which == "widelong":
    for ns in (300000, 1000000, 4000000):
        run("widelong", "board", 64, 148, ns, algo="wide")
    run("widelong", "board", 64, 16384, 300000, algo="gmem")
    run("widelong", "board", 64, 16384, 1000000, algo="gmem")

if which == "wideprof2":
    run("wideprof2", "board", 64, 148, 1000000, algo="wide")

if which == "wident":
    for nt in ("256", "128", "64"):
        os.environ["MCQ_WIDE_THREADS"] = nt
        run("wident" + nt, "board", 30, 2368, 100000, algo="wide")
        run("wident" + nt, "board", 40, 1184, 100000, algo="wide")
        run("wident" + nt, "board", 64, 148, 300000, algo="wide")
        run("wident" + nt, "full_3d", 24, 1184, 100000, algo="wide")

if which == "wideauto":
    os.environ.pop("MCQ_WIDE_THREADS", None)
    for mode, n, nc, ns in (("board", 30, 2368, 100000), ("board", 30, 148, 100000), ("board", 40, 1184, 100000), ("board", 64, 148, 300000),
                            ("full_3d", 24, 1184, 100000), ("full_3d", 40, 1184, 50000), ("board", 22, 4736, 100000)):
        run("wideauto", mode, n, nc, ns)
    run("wideauto-lines", "board", 22, 4736, 100000, algo="lines")

if which == "tablewide":
    for mode in ("board", "full_3d"):
        for n in (12, 16, 20):
            for algo in ("table", "wide"):
                run("tablewide", mode, n, 8192, 50000, algo=algo)

if which == "tablewide2":
    for n in (17, 18, 19, 20):
        for algo in ("table", "wide"):
            run("tablewide2", "full_3d", n, 8192, 50000, algo=algo)
    for n in (20, 21):
        for algo in ("table", "wide"):
            run("tablewide2", "board", n, 16384, 50000, algo=algo)

if which == "wpc":
    for w in (4, 2, 1):
        run("wpc", "full_3d", 12, 20480, 200000, algo="table", warps_per_cta=w)
        run("wpc", "board", 12, 20480, 100000, algo="table", warps_per_cta=w)
