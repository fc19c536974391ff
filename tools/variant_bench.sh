#!/bin/bash
# times prebuilt library variants (variants/*.so, built on the CPU box with -D flags) on the GPU box
cp inquistr_b200/lib/libinqcall.so /tmp/libinqcall.orig.so
for f in ${VARIANTS:-variants/*.so}; do
  cp $f inquistr_b200/lib/libinqcall.so
  timeout 200 python bench.py --scale ${SCALE:-1} --steps ${STEPS:-5} --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$f', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms_rank0'].items()})"
done
cp /tmp/libinqcall.orig.so inquistr_b200/lib/libinqcall.so
