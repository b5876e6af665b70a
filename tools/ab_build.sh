#!/bin/bash
# usage: tools/ab_build.sh name 'sed-expression' [file]   -> builds ab/name.so from a patched copy of the csrc tree
set -e
name=$1; expr=$2; file=${3:-mbm_kernel_v2.cuh}
rm -rf /tmp/ab_$name; mkdir -p /tmp/ab_$name/pkg/csrc /tmp/ab_$name/include ab
cp movingboundaryminerals.jl_b200/csrc/*.cu* /tmp/ab_$name/pkg/csrc/; cp include/mbm_b200.h /tmp/ab_$name/include/
sed -i "$expr" /tmp/ab_$name/pkg/csrc/$file
(cd /tmp/ab_$name/pkg/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-extended-lambda -Xcompiler -fPIC -shared -Xptxas -v -o /root/repo/ab/$name.so mbm_api.cu 2>&1 | grep -E "error|v2ILi256ELi2" -A2 | grep -E "error|spill" | head -3)
