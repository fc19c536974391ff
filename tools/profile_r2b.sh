#!/bin/bash
# GPU box: ncu captures of the cohort kernels and the BGZF inflate kernel (each after the same command exited 0 without ncu)
set -x
Z="python tools/bench_outlier.py --reps 1 --methods zscore"
$Z > gpurun_out/r2_plain_z.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_outlier_zscore -c 1 -f -o gpurun_out/r2_outlier_z $Z > gpurun_out/r2_ncu_z.log 2>&1
D="python tools/bench_outlier.py --reps 1 --methods dbscan --rows 50000"
$D > gpurun_out/r2_plain_d.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_outlier_dbscan -c 1 -f -o gpurun_out/r2_outlier_d $D > gpurun_out/r2_ncu_d.log 2>&1
I="python tools/bench_gpu_inflate.py --scale 0.005 --reps 1"
$I > gpurun_out/r2_plain_i.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_bgzf_inflate -c 1 -f -o gpurun_out/r2_inflate $I > gpurun_out/r2_ncu_i.log 2>&1
ls -la gpurun_out/r2_outlier_z.ncu-rep gpurun_out/r2_outlier_d.ncu-rep gpurun_out/r2_inflate.ncu-rep
