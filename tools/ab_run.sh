#!/bin/bash
# runs bench.py (value only) against every ab/*.so given on the command line, on the same box
for lib in "$@"; do
  out=$(MBM_B200_LIB=$PWD/ab/$lib.so python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1)
  echo "$out" | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', '%.4g' % d['value'], '%.1f ms' % d['ms_per_step'], d['members_failed'])" 2>/dev/null || { echo "$lib FAILED"; echo "$out" | tail -5; }
done
