#!/bin/bash
# runs tools/c5_case.py against every ab/*.so given on the command line, on the same box
for lib in "$@"; do
  echo -n "$lib: "; MBM_B200_LIB=$PWD/ab/$lib.so python tools/c5_case.py 296 20 2>&1 | tail -1
done
