for d in 0 1 2 3; do echo "ZW_DEBUG=$d"; INQ_LIB=$PWD/variants/dbg.so INQ_ZW_DEBUG=$d timeout 200 python tools/bench_outlier.py --reps 2 2>&1 | grep zscore | cut -c1-140; done
echo ROWS; INQ_LIB=$PWD/variants/dbg.so INQ_ZSCORE_ROWS=1 timeout 200 python tools/bench_outlier.py --reps 2 2>&1 | grep zscore | cut -c1-140
