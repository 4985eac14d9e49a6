#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x -k "ib or force or spread or golden or ensemble or step" 2>&1 | tail -2
python bench.py --workload ens1024 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ens1024', round(d['value']), round(d['ms_per_step']*1e3,1), d['stage_us'])"
JAXIB_NO_PDL=1 ncu --set full --clock-control none --import-source on -k regex:"cols_sweep" -s 32 -c 2 -o gpurun_out/prof_r3i python tools/slab_kernel_times.py 8192 8 > gpurun_out/r3i_ncu.log 2>&1
tail -3 gpurun_out/r3i_ncu.log
