#!/bin/bash
# c3 (fixed-interface flux ensemble) speed for every ab/<name>.so given, plus the fixed-interface parity tests for the last one
for lib in "$@"; do echo -n "$lib: "; MBM_B200_LIB=$PWD/ab/$lib.so python tools/c3_case.py 2>&1 | tail -1; done
MBM_B200_LIB=$PWD/ab/${@: -1}.so python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x -k "B1 or B2 or fixed or DCNI or no_interaction or thread_configs or DCGF or DCGMB or edge" 2>&1 | tail -3
