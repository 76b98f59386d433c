cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
rm -f gpurun_out/w38.log
for nt in 64 128 256; do
MCQ_WIDE_THREADS=$nt timeout 600 python - >> gpurun_out/w38.log 2>&1 <<PY
import sys, os
sys.path.insert(0, os.getcwd())
sys.argv=['x','none']
exec(open('scripts/perf_probe.py').read().split("which = sys.argv[1]")[0])
def w(n, nc, ns, mode="board"):
    seeds = torch.arange(nc, dtype=torch.int64).cuda() + 42
    best=0
    for _ in range(2):
        r = eng.run(mode, n, ns, seeds, schedules=SCHEDS[1], history="none", device_buffers=True, want_states=False)
        torch.cuda.synchronize()
        best=max(best, nc*ns/(r.kernel_ms*1e-3))
    print($nt, mode, n, nc, ns, "%.3e"%best, flush=True)
w(64, 296, 1000000); w(64, 2368, 300000); w(48, 592, 1000000); w(40, 1184, 300000)
PY
done
grep -v "^{" gpurun_out/w38.log | grep -v Warning
