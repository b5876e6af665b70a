#!/bin/bash
# full ncu capture (source counters) of the tuned kernel for each ab/<name>.so on a small c4 launch (592 members x 100 steps)
mkdir -p gpurun_out
for lib in "$@"; do
  MBM_AB_NOHASH=1 MBM_B200_LIB=$PWD/ab/$lib.so ncu --set full --clock-control none --import-source on -k regex:mbm_step_kernel_v2 -s 2 -c 1 -f \
    -o gpurun_out/prof_ab_$lib python tools/ab_parity.py 592 100 2 > gpurun_out/ncu_ab_$lib.log 2>&1
  tail -1 gpurun_out/ncu_ab_$lib.log
done
