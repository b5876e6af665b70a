#!/bin/bash
# runs tools/ab_parity.py against every ab/<name>.so given on the command line, on the same box (twice, interleaved)
for rep in 1 2; do for lib in "$@"; do MBM_B200_LIB=$PWD/ab/$lib.so python tools/ab_parity.py 1184 300 16 2>&1 | tail -1; done; done
