#!/bin/bash
# GPU box: the latency-bound panel (config 4) — bench line, then the per-launch device times of the same command
set -x
python -m pytest tests/test_cohort.py -m gpu -x -q > gpurun_out/r2c_cohort_tests.log 2>&1; tail -3 gpurun_out/r2c_cohort_tests.log
python tools/bench_outlier.py --reps 5 > gpurun_out/r2c_outlier.json 2> gpurun_out/r2c_outlier.err; cat gpurun_out/r2c_outlier.json
B="python bench.py --config 4 --steps 200 --warmup 20 --no-cohort"
$B > gpurun_out/r2c_cfg4.json 2> gpurun_out/r2c_cfg4.err && cat gpurun_out/r2c_cfg4.json && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2c_cfg4_launches.csv \
  python bench.py --config 4 --steps 3 --warmup 3 --no-cohort --no-graph > gpurun_out/r2c_cfg4_ncu.log 2>&1
tail -2 gpurun_out/r2c_cfg4_ncu.log
