#!/bin/bash
# compile-time sweep of the scan kernel's pipeline shape on the GPU box (timing only)
cp inquistr_b200/lib/libinqcall.so /tmp/libinqcall.orig.so
for cfg in ${CFGS:-"16 16 3 2" "32 16 2 1" "32 8 3 2" "32 12 2 2" "32 16 3 1" "32 10 2 2"}; do
  set -- $cfg
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC \
     -DINQ_LANE_WORDS=$1 -DINQ_SCAN_WARPS=$2 -DINQ_WARP_STAGES=$3 -DINQ_SCAN_MIN_CTAS=$4 -o inquistr_b200/lib/libinqcall.so inquistr_b200/csrc/inq_capi.cu inquistr_b200/csrc/inq_cohort_capi.cu 2>/dev/null
  timeout 200 python bench.py --scale ${SCALE:-0.5} --steps 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('lane_words $1 warps/cta $2 stages $3 minctas $4 ms_cigar', round(d['stage_ms_rank0']['ms_cigar'],3), 'GB/s', round(d['roofline']['streamed_GBps']), 'fixup', round(d['stage_ms_rank0']['ms_fixup'],3))"
  timeout 100 python tools/sanitize_small.py 2>&1 | tail -1
done
cp /tmp/libinqcall.orig.so inquistr_b200/lib/libinqcall.so
