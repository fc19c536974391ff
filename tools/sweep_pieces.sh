#!/bin/bash
# GPU box: how many median chunks per pass (each: median kernel + deep-locus kernel + push kernel) at full size and at 1/8
for s in 1.0 0.125; do
for p in 2 4 8 12 16; do
for m in $( [ $s = 1.0 ] && echo 32768 || echo "16384 32768 65536" ); do
  python bench.py --scale $s --set median_pieces=$p --set min_piece=$m --steps 40 --warmup 5 --no-cohort --no-e2e --no-cpu-baseline --bam-scale 0 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'scale': $s, 'pieces': $p, 'min_piece': $m, 'ms_per_step': round(d['ms_per_step'], 4), 'device_ms': round(d['device_ms_per_step'], 4), 'chunks': d['pipeline'].get('median_chunks'), 'median': round(d['stage_ms_rank0']['ms_median'], 3), 'd2h': round(d['stage_ms_rank0']['ms_d2h'], 3)}))"
done; done; done | tee gpurun_out/r2d_pieces.jsonl
