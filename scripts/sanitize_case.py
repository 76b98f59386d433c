"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): both kernels, both modes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import monte_carlo_collective_b200 as mcq  # noqa: E402
from monte_carlo_collective_b200 import schedules  # noqa: E402

eng = mcq.Engine(0)
ns = 600
betas = schedules.beta_table({"type": "linear_annealing", "beta_start": 0.5, "beta_end": 3.0}, ns)
for mode in ("board", "full_3d"):
    for kw in (dict(algo="table"), dict(algo="table", lanes_per_chain=16), dict(algo="lines", lanes_per_chain=8),
               dict(algo="lines", lanes_per_chain=32)):
        r = eng.run(mode, 7, ns, np.arange(13, dtype=np.uint64), betas, history="stats", n_bins=10, accept_bits=True,
                    chunk_steps=256, early_stop_patience=200 if mode == "board" else None, **kw)
        assert (eng.energy(mode, 7, r.final_state) == r.final_energy).all()
        print(mode, kw, int(r.best_energy.min()))
print("sanitize case ok")
