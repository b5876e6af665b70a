#!/bin/bash
# full GPU verification of the current tree: parity tests, smoke, both bench arms, ncu launch list + one full capture.
# usage: tools/gpu_verify.sh [tag]   (files go to gpurun_out/<tag>_*)
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.log 2>&1
python bench.py > gpurun_out/${tag}_bench_full.log 2> gpurun_out/${tag}_bench_full.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench_full.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mbm_step_kernel_v2 -c 1 -f -o gpurun_out/prof_${tag}_v2_head \
  python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/${tag}_ncu_full.log 2>&1
tail -3 gpurun_out/${tag}_pytest_gpu.log; tail -2 gpurun_out/${tag}_smoke.log; tail -1 gpurun_out/${tag}_bench_full.log | cut -c1-400
python tools/measure_configs.py > gpurun_out/${tag}_configs.log 2>&1; tail -4 gpurun_out/${tag}_configs.log | cut -c1-250
