cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
cat > /tmp/w.py <<'PY'
import os, sys
sys.path.insert(0, os.environ["GRAFT_REPO_ROOT"])
import numpy as np, torch
import monte_carlo_collective_b200 as mcq
eng = mcq.Engine(0)
LIN = {"type": "linear_annealing", "beta_start": 1.0, "beta_end": 3.0}
def run(n, nc, ns, mode="board", **kw):
    seeds = torch.arange(nc, dtype=torch.int64).cuda() + 42
    best = 0
    for _ in range(2):
        r = eng.run(mode, n, ns, seeds, schedules=LIN, history="none", device_buffers=True, want_states=False, **kw)
        torch.cuda.synchronize()
        best = max(best, nc * ns / (r.kernel_ms * 1e-3))
    print(os.environ.get("MCQ_WIDE_THREADS", "auto"), mode, n, nc, ns, kw, "%.3e" % best, flush=True)
for n, nc in ((22, 4736), (30, 2368), (40, 1184)):
    run(n, nc, 100000)
run(24, 1184, 100000, mode="full_3d")
run(64, 296, 300000)
PY
for t in auto 32 64 128 256; do
  if [ $t = auto ]; then python /tmp/w.py; else MCQ_WIDE_THREADS=$t python /tmp/w.py; fi
done > gpurun_out/w16.log 2>&1
for m in 6 8 11 13 16; do python - <<PY >> gpurun_out/w16.log 2>&1
import os, sys
sys.path.insert(0, os.environ["GRAFT_REPO_ROOT"])
import torch, monte_carlo_collective_b200 as mcq
eng = mcq.Engine(0)
seeds = torch.arange(4736, dtype=torch.int64).cuda() + 42
for _ in range(2):
    r = eng.run("board", 22, 100000, seeds, schedules={"type": "linear_annealing", "beta_start": 1.0, "beta_end": 3.0}, history="none", device_buffers=True, want_states=False, algo="table")
    torch.cuda.synchronize()
print("N=22 table kernel", "%.3e" % (4736 * 100000 / (r.kernel_ms * 1e-3)))
PY
break; done
cat gpurun_out/w16.log
