"""Small cases of the CTA-per-chain kernel's multi-commit rounds for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import monte_carlo_collective_b200 as mcq  # noqa: E402

eng = mcq.Engine(0)
ns = 400
sched = {"type": "linear_annealing", "beta_start": 0.4, "beta_end": 3.0}
for mode, n, warps in (("board", 40, 2), ("board", 26, 1), ("board", 33, 4), ("full_3d", 20, 2), ("full_3d", 12, 1)):
    r = eng.run(mode, n, ns, np.arange(3, dtype=np.uint64) + 7, schedules=sched, history="full", n_bins=7, accept_bits=True,
                algo="wide", warps_per_cta=warps)
    assert (eng.energy(mode, n, r.final_state) == r.final_energy).all()
    assert (eng.energy(mode, n, r.best_state) == r.best_energy).all()
    print(mode, n, warps, int(r.best_energy.min()), int(r.n_accepted.sum()))
print("sanitize wide ok")
