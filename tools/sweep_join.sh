#!/bin/bash
# GPU box: warp-cooperative join (1) vs two scalar binary searches per read (0)
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for cfg in "3 1.0" "3 0.125" "2 1.0" "4 1.0" "5 0.25"; do
set -- $cfg
for v in 1 0; do
  python bench.py --config $1 --scale $2 --set join_coop=$v --steps 50 --warmup 5 --no-cohort --no-e2e --no-cpu-baseline --bam-scale 0 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'config': $1, 'scale': $2, 'join_coop': $v, 'ms_per_step': round(d['ms_per_step'], 4), 'device_ms': round(d['device_ms_per_step'], 4), 'frac': round(d['roofline']['frac'], 3), 'parity': (d.get('parity') or {}).get('bit_exact_vs_oracle'), 'stage': {k: round(v, 3) for k, v in d['stage_ms_rank0'].items()}}))"
done; done | tee gpurun_out/r2h_join.jsonl
