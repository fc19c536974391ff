#!/bin/bash
# GPU box: grid of k_push_results (it shares the SMs with the medians of the next chunk)
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for s in 1.0 0.125; do
for v in 2 8 32 64 148; do
  python bench.py --scale $s --set push_ctas=$v --steps 50 --warmup 5 --no-cohort --no-e2e --no-cpu-baseline --bam-scale 0 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'scale': $s, 'push_ctas': $v, 'ms_per_step': round(d['ms_per_step'], 4), 'device_ms': round(d['device_ms_per_step'], 4), 'median': round(d['stage_ms_rank0']['ms_median'], 3), 'd2h': round(d['stage_ms_rank0']['ms_d2h'], 3)}))"
done; done | tee gpurun_out/r2g_push_ctas.jsonl
