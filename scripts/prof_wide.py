"""One launch of the CTA-per-chain kernel at C5's board size, for ncu (development tool).
   python scripts/prof_wide.py [chains] [steps]   -- N=64 board, linear 1->3, no history"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
import monte_carlo_collective_b200 as mcq  # noqa: E402
import torch  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 296
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 300000
eng = mcq.Engine(0)
seeds = torch.arange(nc, dtype=torch.int64).cuda() + 42
for _ in range(2):
    r = eng.run("board", 64, ns, seeds, schedules={"type": "linear_annealing", "beta_start": 1.0, "beta_end": 3.0}, history="none",
                device_buffers=True, want_states=False)
    torch.cuda.synchronize()
print("wide N=64", nc, "chains pps %.3e" % (nc * ns / (r.kernel_ms * 1e-3)), "acc %.4f" % (float(r.n_accepted.double().mean()) / ns), "ms", r.kernel_ms)
