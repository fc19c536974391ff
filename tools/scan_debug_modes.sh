#!/bin/bash
# timing experiments for k_cigar_scan (INQ_SCAN_DEBUG bit mask; results are wrong when != 0)
# bit0 no event emission, bit1 no look-back, bit2 no read-start staging, bit3 no phase-A math,
# bit4 no publish / ev_off stores (ONLY together with bit1: other CTAs would spin forever otherwise)
for m in ${MODES:-0 1 2 3 6 7 22 23 31}; do
  INQ_SCAN_DEBUG=$m timeout 90 python bench.py --scale ${SCALE:-0.25} --steps 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('mode', $m, 'ms_cigar', round(d['stage_ms_rank0']['ms_cigar'],3), 'GB/s', round(d['roofline']['streamed_GBps']))
except Exception as e: print('mode', $m, 'failed')"
done
