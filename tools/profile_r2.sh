#!/bin/bash
# GPU box: ncu evidence of round 2 (each capture after the same command has exited 0 without ncu). Outputs in gpurun_out/.
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph"
$B > gpurun_out/r2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_l.log 2>&1
B1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph"
$B1 > gpurun_out/r2_plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cigar_scan -s 3 -c 1 -f -o gpurun_out/r2_scan $B1 > gpurun_out/r2_ncu_s.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base function -k 'regex:^k_pair_eval$|^k_locus_median$|^k_join_ranges$|^k_exclusive_scan2$' -s 36 -c 12 -f -o gpurun_out/r2_secondary $B1 > gpurun_out/r2_ncu_p.log 2>&1
Z="python tools/bench_outlier.py --reps 1"
$Z > gpurun_out/r2_plain_z.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_outlier -c 2 -f -o gpurun_out/r2_outlier $Z > gpurun_out/r2_ncu_z.log 2>&1
ls -la gpurun_out/*.ncu-rep
