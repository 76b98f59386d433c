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
run(64, 296, 3000000)
run(64, 2368, 1000000)
run(48, 592, 1000000)
run(32, 2368, 300000)
PY
for t in 32 64 128; do MCQ_WIDE_THREADS=$t python /tmp/w.py; done > gpurun_out/w17.log 2>&1
cat gpurun_out/w17.log
