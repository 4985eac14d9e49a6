#!/bin/bash
# Round-end single-GPU record: tests, default bench, the other BASELINE workloads, ncu launch list + full capture.
cd "$(dirname "$0")/.."
T=${1:-r3g}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/${T}_tests.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
for w in flap600 bearing300 channel2048 ens1024; do
  timeout 400 python bench.py --workload $w --steps 50 --warmup 5 > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${T}_bench_$w.json").read().strip().splitlines()[-1])
    print("$w", d["value"], d["ms_per_step"], d["step_roofline"]["frac"], d["config"].get("finite"), d["stage_us"])
except Exception as e:
    print("$w ERR", e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 20 --warmup 3 --no-slab --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ring|interp_force|spread_bodies|rows_fwd|cols_sweep|rows_inv" -s 60 -c 6 -o gpurun_out/prof_${T} python tools/stage_times.py ib2048 3 > gpurun_out/${T}_ncu_full.log 2>&1
tail -2 gpurun_out/${T}_ncu_full.log
