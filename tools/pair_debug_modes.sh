#!/bin/bash
# timing experiments for k_pair_eval (INQ_PAIR_DEBUG: 1 no atomic, 2 no store, 4 no event loads); results wrong when != 0
for m in 0 1 2 4 3 7; do
  INQ_PAIR_DEBUG=$m timeout 120 python bench.py --scale ${SCALE:-0.5} --steps 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('mode', $m, 'ms_pairs', round(d['stage_ms_rank0']['ms_pairs'],3))"
done
