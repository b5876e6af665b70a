#!/bin/bash
# e2e (host buffers through the C ABI) of the default bench workload for the library in ab/$1.so; extra env in $2..
lib=$1; shift
env "$@" MBM_B200_LIB=$PWD/ab/$lib.so python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib $*', 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], '%.1f ms' % d['e2e']['ms_per_step'])"
