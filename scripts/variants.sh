#!/bin/bash
# time each prebuilt library variant on the C2-like workload (development tool)
for lib in monte_carlo_collective_b200/variants/*.so monte_carlo_collective_b200/libmcq.so; do
  echo "== $lib"
  MCQ_LIB_PATH=$PWD/$lib python scripts/sweep.py ${1:-quick} 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], d['mode'], d['n'], d['steps'], '%.3e' % d['pps'], 'acc=%.3f' % d['acc'])"
done
