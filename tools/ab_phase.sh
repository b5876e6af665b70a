#!/bin/bash
# phase clocks (libraries built with -DMBM_PHASE_CLOCKS): cycles per step between the barriers of one CTA, for block $MBM_PHASE_BLOCK
for lib in "$@"; do echo "== $lib"; MBM_B200_LIB=$PWD/ab/$lib.so python tools/ab_parity.py 592 200 2 2>&1 | grep -E "phase cycles|node-steps" | tail -3; done
