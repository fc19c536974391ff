#!/bin/bash
# GPU box: results through k_push_results (1) or three copy-engine operations per chunk (0), at full size and at 1/8
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for s in 1.0 0.125; do
for v in 1 0 1 0; do
  python bench.py --scale $s --set push_kernel=$v --steps 50 --warmup 5 --no-cohort --no-e2e --no-cpu-baseline --bam-scale 0 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'scale': $s, 'push_kernel': $v, 'ms_per_step': round(d['ms_per_step'], 4), 'device_ms': round(d['device_ms_per_step'], 4), 'launches': d['gpu_launches'], 'stage': {k: round(v, 3) for k, v in d['stage_ms_rank0'].items()}}))"
done; done | tee gpurun_out/r2d_push_kernel.jsonl
