#!/bin/bash
# GPU box: timing experiments on k_outlier_zscore_warp (variants/dbg.so = -DINQ_TIMING_EXPERIMENTS; results are wrong for bits 1-8)
#   bit 1 skip the flag pass, 2 skip the sums, 4 approximate threshold (no walk), 8 count hits only
for c in 3 30; do for d in 0 1 3 4 8 12; do echo "cutoff=$c ZW_DEBUG=$d"; INQ_LIB=$PWD/variants/dbg.so INQ_ZW_DEBUG=$d timeout 200 python tools/bench_outlier.py --reps 2 --methods zscore --cutoff $c 2>&1 | grep zscore | cut -c1-150; done; done
echo ROWS; INQ_LIB=$PWD/variants/dbg.so INQ_ZSCORE_ROWS=1 timeout 200 python tools/bench_outlier.py --reps 2 --methods zscore 2>&1 | grep zscore | cut -c1-150
