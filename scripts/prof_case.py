"""One launch of the conflict-table kernel in a fixed regime, for ncu (development tool).
   python scripts/prof_case.py cold|hot|mid [chains] [steps]   -- constant beta 5.0 / 1.0 / 2.0, stats mode, N=12 full_3d"""
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

which = sys.argv[1] if len(sys.argv) > 1 else "cold"
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 60000
beta = {"cold": 5.0, "hot": 1.0, "mid": 2.0}[which]
mode = os.environ.get("MCQ_MODE", "full_3d")
eng = mcq.Engine(0)
seeds = torch.arange(nc, dtype=torch.int64).cuda() + 42
for _ in range(2):
    r = eng.run(mode, 12, ns, seeds, schedules={"type": "constant", "beta_const": beta}, history="stats", n_bins=100,
                device_buffers=True, want_states=False, init_mode=os.environ.get("MCQ_INIT", "random"))
    torch.cuda.synchronize()
print(which, "pps %.3e" % (nc * ns / (r.kernel_ms * 1e-3)), "acc %.4f" % (float(r.n_accepted.double().mean()) / ns), "ms", r.kernel_ms)
