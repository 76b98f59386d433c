"""Run BASELINE.json configs C1, C3, C4 and a bounded sample of C5 on one GPU; write profiles/r1_configs.json.

Records throughput and the qualitative anchors of BASELINE.md (min energy vs N per initialisation, best beta range).
"""
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
from monte_carlo_collective_b200 import drivers, schedules  # noqa: E402

eng = mcq.default_engine()
out = {"device": eng.device_name}
LIN = {"type": "linear_annealing", "beta_start": 1.0, "beta_end": 3.0}

# ---- C1: N=8 board, 10 runs x 1e5 steps through the drop-in run_experiment ----
t0 = time.time()
hist, best, times, acc, rej, s2b = mcq.run_experiment(8, 100000, "random", None, 10, base_seed=42, schedule_params=LIN,
                                                      mcmc_type="board", early_stop_patience=None)
out["C1"] = {"best_energies": best, "steps_to_best": s2b, "accept_counts": [len(a) for a in acc],
             "wall_s": time.time() - t0, "reference_best_energies": [63, 56, 56, 56, 65, 61, 68, 56, 66, 62]}

# ---- C3: min energy vs N, three initialisations (report Fig. 5 / 6), 256 replicas x 1e6 steps ----
Ns = list(range(3, 16))
t0 = time.time()
res = drivers.measure_min_energy_vs_N(Ns, 1000000, None, schedule_params=LIN, init_modes=["random", "klarner", "latin"],
                                      n_runs=256, base_seed=100, verbose=False, plot=True, mcmc_type="board",
                                      early_stop_patience=None, results_dir=os.path.join(ROOT, "profiles", "r1_results"))
dt = time.time() - t0
out["C3"] = {"Ns": Ns, "n_runs": 256, "n_steps": 1000000, "wall_s": dt, "proposals_per_s": 3 * len(Ns) * 256 * 1e6 / dt,
             "mean_min_energy": {k: v["mean_min_energies"].tolist() for k, v in res["results"].items()},
             "min_min_energy": {k: [int(a.min()) for a in v["all_min_energies"]] for k, v in res["results"].items()},
             "mean_steps_to_best": {k: v["mean_steps_to_best"].tolist() for k, v in res["results"].items()}}

# ---- C4: N=20 board, exponential annealing, 64 (beta_start, beta_end) pairs x 1024 replicas x 1e6 steps ----
starts = [0.1, 0.25, 0.5, 0.75, 1.0, 1.5, 2.0, 3.0]
ends = [2.0, 3.0, 4.0, 5.0, 6.0, 8.0, 10.0, 20.0]
pairs = [[a, b] for a in starts for b in ends]
t0 = time.time()
r4 = drivers.run_beta_start_end_pairs(N=20, n_steps=1000000, beta_start_ends=pairs, annealing_type="exponential_annealing",
                                      n_runs=1024, base_seed=42, verbose=False, plot=False, mcmc_type="board",
                                      early_stop_patience=None, history="stats")
dt = time.time() - t0
mean_best = {k: float(np.mean(v)) for k, v in r4["all_best_energies"].items()}
out["C4"] = {"wall_s": dt, "proposals_per_s": 64 * 1024 * 1e6 / dt, "mean_best_energy": mean_best,
             "best_pair": min(mean_best, key=mean_best.get), "final_mean_energy": {k: float(v[-1]) for k, v in r4["mean_energy"].items()}}

# ---- C5 (bounded sample): N=64 board, 65536 replicas, global-memory line counters, first 2e4 of the 1e7 steps ----
ns = 20000
betas = schedules.beta_table(LIN, 10000000)[:ns]        # the first steps of the 1e7-step schedule
r5 = eng.run("board", 64, ns, np.arange(65536, dtype=np.uint64), betas, history="none", want_states=False)
out["C5_sample"] = {"chains": 65536, "n_steps": ns, "kernel_ms": r5.kernel_ms, "proposals_per_s": 65536 * ns / (r5.kernel_ms * 1e-3),
                    "mean_best_energy": float(r5.best_energy.mean()), "acceptance": float(r5.n_accepted.mean()) / ns,
                    "note": "N=64 needs 121 KB of line counters per chain: one thread per chain, counters in global memory (8.3 GB); "
                            "bounded sample of the hot start of C5 (65536 chains x 1e7 steps = 6.6e11 proposals)"}
# ---- C5 in full (MCQ_FULL_C5=1): 65536 chains x 1e7 steps = 6.55e11 proposals, launches of 5e5 steps ----
if os.environ.get("MCQ_FULL_C5") == "1":
    ns = 10000000
    betas = schedules.beta_table(LIN, ns)
    t0 = time.time()
    r5 = eng.run("board", 64, ns, np.arange(65536, dtype=np.uint64), betas, history="none", want_states=False, chunk_steps=500000)
    out["C5_full"] = {"chains": 65536, "n_steps": ns, "kernel_ms": r5.kernel_ms, "wall_s": time.time() - t0,
                      "proposals_per_s": 65536.0 * ns / (r5.kernel_ms * 1e-3), "gpu_launches": int(r5.gpu_launches),
                      "mean_best_energy": float(r5.best_energy.mean()), "min_best_energy": int(r5.best_energy.min()),
                      "mean_final_energy": float(r5.final_energy.mean()), "mean_initial_energy": float(r5.initial_energy.mean()),
                      "acceptance": float(r5.n_accepted.mean()) / ns}
path = os.path.join(ROOT, "profiles", "r1_configs.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps({k: (v if k == "device" else {kk: vv for kk, vv in v.items() if kk in ("wall_s", "proposals_per_s", "best_pair", "best_energies", "mean_best_energy", "min_best_energy", "acceptance", "kernel_ms")}) for k, v in out.items()}))
print("mean min energy (random):", [round(x, 1) for x in out["C3"]["mean_min_energy"]["random"]])
print("mean min energy (klarner):", [round(x, 1) for x in out["C3"]["mean_min_energy"]["klarner"]])
print("mean min energy (latin):", [round(x, 1) for x in out["C3"]["mean_min_energy"]["latin"]])
