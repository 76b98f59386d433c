"""BASELINE config C3: measure_min_energy_vs_N, Ns = 3..15 x {random, klarner, latin}, replicas sharded over every visible
GPU -- through the reference-shaped driver (drivers.measure_min_energy_vs_N -> multi.DevicePool.run_many: the 39
independent problems are dealt over the devices and issued from several host threads per device, one stream each).
usage: python scripts/run_c3.py [n_runs] [n_steps] [out.json]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
from monte_carlo_collective_b200 import drivers, multi  # noqa: E402

n_runs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_steps = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "c3.json")
Ns = list(range(3, 16))
inits = ["random", "klarner", "latin"]                      # config.yaml:24
sp = {"type": "linear_annealing", "beta_start": 1.0, "beta_end": 3.0}
devices = multi.visible_devices()
out = {"devices": devices, "n_runs": n_runs, "n_steps": n_steps, "Ns": Ns, "init_modes": inits}
for mcmc_type in ("board", "full_3d"):
    best = None
    for rep in range(2):                                     # the first pass creates contexts, neighbour tables, ...
        t0 = time.perf_counter()
        res = drivers.measure_min_energy_vs_N(Ns, n_steps, None, schedule_params=sp, init_modes=inits, n_runs=n_runs, base_seed=100,
                                              verbose=False, plot=False, mcmc_type=mcmc_type, early_stop_patience=None, workers=8)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    proposals = len(Ns) * len(inits) * n_runs * n_steps
    out[mcmc_type] = {"seconds": best, "proposals": proposals, "proposals_per_s": proposals / best,
                      "mean_min_energy": {k: [float(x) for x in v["mean_min_energies"]] for k, v in res["results"].items()}}
    print(mcmc_type, "%.2f s" % best, "%.3e proposals/s" % (proposals / best), flush=True)
json.dump(out, open(out_path, "w"), indent=1)
