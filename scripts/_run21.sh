cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/t21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t21.log
rm -f gpurun_out/probe.jsonl gpurun_out/p21.log
for lib in monte_carlo_collective_b200/libmcq.so monte_carlo_collective_b200/variants/libmcq_board4x32.so; do
  echo "== $lib" >> gpurun_out/p21.log
  MCQ_LIB_PATH=$PWD/$lib timeout 600 python scripts/perf_probe.py board 1000000 >> gpurun_out/p21.log 2>&1
  MCQ_LIB_PATH=$PWD/$lib timeout 600 python - >> gpurun_out/p21.log 2>&1 <<'PY'
import sys; sys.argv=['x','none']
import runpy, os
sys.path.insert(0, os.getcwd())
exec(open('scripts/perf_probe.py').read().split("which = sys.argv[1]")[0])
whole(1000000, mode="board", n=20, reps=2048)
whole(1000000, mode="board", n=16, reps=2048)
PY
done
timeout 600 python scripts/run_c3.py 1024 1000000 gpurun_out/c3_21.json > gpurun_out/c3_21.log 2>&1
tail -3 gpurun_out/t21.log; grep -h "^==\|^{" gpurun_out/p21.log | cut -c1-180; cat gpurun_out/c3_21.log
