#!/bin/bash
# one full ncu capture of the tuned kernel on a small c4 launch (592 members x 100 steps): plain run first, then under ncu
mkdir -p gpurun_out
python tools/ab_parity.py 592 100 2 > gpurun_out/ncu_small_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mbm_step_kernel_v2 -s 2 -c 1 -f -o gpurun_out/prof_v2_small \
  python tools/ab_parity.py 592 100 2 > gpurun_out/ncu_small.log 2>&1
tail -2 gpurun_out/ncu_small_plain.log; tail -3 gpurun_out/ncu_small.log
