"""Throughput probe of the conflict-table kernel on the C2 workload (development tool).

  python scripts/perf_probe.py whole [steps]      all five schedules, one call, stats mode: proposals/s
  python scripts/perf_probe.py phases [steps]     per schedule, 8 segments: time and acceptance per segment,
                                                   for 32 and 16 lanes per chain (size-generic kernels)
Writes JSON lines to gpurun_out/probe.jsonl (tagged with $MCQ_TAG / $MCQ_LIB_PATH).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

if not os.environ.get("MCQ_LIB_PATH"):
    ge.build()
import monte_carlo_collective_b200 as mcq  # noqa: E402
import torch  # noqa: E402

SCHEDS = [
    {"type": "constant", "beta_const": 5.0},
    {"type": "linear_annealing", "beta_start": 1.0, "beta_end": 3.0},
    {"type": "exponential_annealing", "beta_start": 1.0, "beta_end": 3.0},
    {"type": "logarithmic_annealing", "beta_start": 1.0, "beta_end": 3.0},
    {"type": "sinusoidal_annealing", "beta_start": 1.0, "beta_end": 3.0},
]
TAG = os.environ.get("MCQ_TAG", os.path.basename(os.environ.get("MCQ_LIB_PATH", "default")))
eng = mcq.Engine(0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "probe.jsonl"), "a")


def emit(d):
    d = dict(tag=TAG, **d)
    print(json.dumps(d), flush=True)
    out.write(json.dumps(d) + "\n")
    out.flush()


def whole(ns, mode="full_3d", n=12, reps=4096, history="stats", **kw):
    seeds = torch.from_numpy(np.tile(np.arange(reps, dtype=np.int64) + 42, len(SCHEDS))).cuda()
    groups = torch.from_numpy(np.repeat(np.arange(len(SCHEDS), dtype=np.int32), reps)).cuda()
    best = None
    for _ in range(3):
        r = eng.run(mode, n, ns, seeds, schedules=SCHEDS, groups=groups, history=history, n_bins=100, device_buffers=True, **kw)
        torch.cuda.synchronize()
        pps = len(seeds) * ns / (r.kernel_ms * 1e-3)
        if best is None or pps > best["pps"]:
            best = dict(what="whole", mode=mode, n=n, steps=ns, history=history, kernel_ms=r.kernel_ms, pps=pps,
                        acc=float(r.n_accepted.double().mean()) / ns, launches=r.gpu_launches,
                        near=int(r.n_near_threshold.sum()), flips=int(r.n_fp32_flips.sum()), **kw)
    emit(best)


def phases(ns, lanes, mode="full_3d", n=12, reps=4096, segs=8):
    for g, sp in enumerate(SCHEDS):
        seeds = torch.arange(reps, dtype=torch.int64).cuda() + 42
        res, prev_acc, t = None, 0.0, 0
        seg = ns // segs // 32 * 32
        rows = []
        while t < ns:
            stop = min(ns, t + seg) if t + 2 * seg <= ns else ns
            res = eng.run(mode, n, ns, seeds, schedules=sp, history="none", device_buffers=True, algo="table",
                          lanes_per_chain=lanes, resume=res, stop_step=stop if stop < ns else None, want_states=True)
            torch.cuda.synchronize()
            acc = float(res.n_accepted.double().sum())
            rows.append(dict(t0=t, t1=stop, ms=round(res.kernel_ms, 3), p=round((acc - prev_acc) / (reps * (stop - t)), 4),
                             pps=reps * (stop - t) / (res.kernel_ms * 1e-3)))
            prev_acc, t = acc, stop
        emit(dict(what="phases", sched=sp["type"], lanes=lanes, steps=ns, total_ms=round(sum(r["ms"] for r in rows), 2), segs=rows))


def wide(n, nc, ns, mode="board", **kw):
    seeds = torch.arange(nc, dtype=torch.int64).cuda() + 42
    best = None
    for _ in range(2):
        r = eng.run(mode, n, ns, seeds, schedules=SCHEDS[1], history="none", device_buffers=True, want_states=False, **kw)
        torch.cuda.synchronize()
        pps = nc * ns / (r.kernel_ms * 1e-3)
        if best is None or pps > best["pps"]:
            best = dict(what="wide", mode=mode, n=n, chains=nc, steps=ns, kernel_ms=r.kernel_ms, pps=pps,
                        acc=float(r.n_accepted.double().mean()) / ns, mean_best=float(r.best_energy.double().mean()), **kw)
    emit(best)


which = sys.argv[1] if len(sys.argv) > 1 else "whole"
ns = int(float(sys.argv[2])) if len(sys.argv) > 2 else 200000
if which == "whole":
    whole(ns)
    whole(ns, history="none")
elif which == "board":
    whole(ns, mode="board")
elif which == "wide":
    for nc, steps in ((148, 100000), (296, 100000), (296, 1000000), (1184, 300000), (4736, 300000), (65536, 100000)):
        wide(64, nc, steps)
    wide(30, 2368, 100000)
    wide(22, 4736, 100000)
    wide(24, 1184, 100000, mode="full_3d")
    wide(40, 1184, 100000)
elif which == "phases":
    for lanes in (32, 16):
        phases(ns, lanes)
