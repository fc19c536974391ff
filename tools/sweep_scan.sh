#!/bin/bash
# compile-time sweep of the scan kernel's pipeline shape on the GPU box (timing only)
set -e
cp inquistr_b200/lib/libinqcall.so /tmp/libinqcall.orig.so
for cfg in "8 3 4" "8 4 3" "16 3 2" "8 6 2" "4 6 4" "16 2 3" "4 3 8"; do
  set -- $cfg
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC \
     -DINQ_SCAN_WARPS=$1 -DINQ_WARP_STAGES=$2 -DINQ_SCAN_MIN_CTAS=$3 -o inquistr_b200/lib/libinqcall.so inquistr_b200/csrc/inq_capi.cu 2>/dev/null
  timeout 120 python bench.py --scale ${SCALE:-0.5} --steps 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('warps/cta $1 stages $2 minctas $3 ms_cigar', round(d['stage_ms_rank0']['ms_cigar'],3), 'GB/s', round(d['roofline']['streamed_GBps']))"
done
cp /tmp/libinqcall.orig.so inquistr_b200/lib/libinqcall.so
