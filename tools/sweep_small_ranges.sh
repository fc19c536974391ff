#!/bin/bash
# GPU box: does the range pipeline pay at the per-rank size of an 8-GPU run (1/8 of config 3)?
for s in 0.125 0.25; do
for k in 1 2 3 4 6 8; do
  python bench.py --scale $s --ranges $k --steps 50 --warmup 5 --no-cohort --no-e2e --no-cpu-baseline --bam-scale 0 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'scale': $s, 'ranges': $k, 'ms_per_step': round(d['ms_per_step'], 4), 'device_ms': round(d['device_ms_per_step'], 4), 'pipeline': d['pipeline'].get('ranges'), 'stage': {k: round(v, 3) for k, v in d['stage_ms_rank0'].items()}}))"
done; done | tee gpurun_out/r2d_small_ranges.jsonl
