#!/bin/bash
# GPU box: final validation of the round — build check, smoke, the whole GPU suite, the bench line, the two side
# configurations, and the launch list of the same bench command (after it exited 0 without ncu).
set -x
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_gpu_tests.log 2>&1; tail -3 gpurun_out/r2f_gpu_tests.log
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 400 gpurun_out/r2f_bench.err
for c in 2 4; do
  python bench.py --config $c --steps 200 --warmup 20 --no-cohort --bam-scale 0 --no-cpu-baseline > gpurun_out/r2f_cfg$c.json 2> gpurun_out/r2f_cfg$c.err
done
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --no-cohort --bam-scale 0"
$B > gpurun_out/r2f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches.csv $B > gpurun_out/r2f_ncu_l.log 2>&1
python - <<'PY'
import json
for f in ("r2f_bench", "r2f_cfg2", "r2f_cfg4"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["device_ms_per_step"], d["roofline"]["frac"], d["roofline"].get("pipeline_frac"), d["gpu_launches"], d.get("e2e", {}).get("value") if d.get("e2e") else None, d.get("parity"))
    except Exception as e:
        print(f, "failed", e)
PY
