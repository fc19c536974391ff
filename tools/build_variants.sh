#!/bin/bash
# CPU box: compile libinqcall.so variants into variants/ (git-ignored; they travel to the GPU box with gpurun).
# usage: tools/build_variants.sh "name:-DINQ_SCAN_WARPS=20 -DINQ_PAIR_WARPS=4" ...
set -e
mkdir -p variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC $flags \
     -o variants/$name.so inquistr_b200/csrc/inq_capi.cu inquistr_b200/csrc/inq_cohort_capi.cu inquistr_b200/csrc/inq_inflate_capi.cu &
done
wait
ls -la variants/
