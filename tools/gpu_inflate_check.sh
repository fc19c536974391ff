#!/bin/bash
# GPU box: `inquistr-b200 call` end to end on one synthetic BAM with and without the GPU inflate engine
for s in ${SCALES:-0.1}; do
for g in ${MODES:-0 1 0 1}; do
  echo "scale $s INQ_GPU_INFLATE=$g"; INQ_GPU_INFLATE=$g timeout 900 python tools/bench_bam.py --scale $s --with-seq -t 16 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['cli_stats']; print({k:d[k] for k in ('cli_wall_s','inflate_GBps_during_scan','tsv_identical_to_oracle')}, {k:v for k,v in c.items() if k.startswith('s_')}, c.get('bytes_inflated_on_gpu'))"
done; done
