#!/bin/bash
# GPU box: what do the per-stage event records cost inside the replayed graph?
for s in 1.0 0.125; do
for v in 1 0 1 0; do
  python bench.py --scale $s --set timing=$v --steps 50 --warmup 5 --no-cohort --no-e2e --no-cpu-baseline --bam-scale 0 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'scale': $s, 'timing': $v, 'ms_per_step': round(d['ms_per_step'], 4), 'device_ms': round(d['device_ms_per_step'], 4), 'stage': {k: round(v, 3) for k, v in d['stage_ms_rank0'].items()}}))"
done; done | tee gpurun_out/r2d_timing.jsonl
