#!/bin/bash
# usage: tools/ab_flags.sh name [extra nvcc flags...]   -> builds ab/name.so from the current csrc tree with extra -D flags
set -e
name=$1; shift
mkdir -p ab
cd movingboundaryminerals.jl_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-extended-lambda -Xcompiler -fPIC -shared \
  -Xptxas -v "$@" -o /root/repo/ab/$name.so mbm_api.cu 2>&1 | grep -E "error|warning: v|v2ILi256ELi2" -A2 | grep -E "error|warning|spill|registers" | head -8
