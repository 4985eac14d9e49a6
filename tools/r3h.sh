#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_slab.py tests/test_dropin_api.py -m gpu -q -x 2>&1 | tail -3
python tools/step_time.py ib2048 30 | tail -1
python tools/step_time.py slab8192 20 | tail -1
JAXIB_K6_MARCH=0 python tools/step_time.py slab8192 20 | tail -1
JAXIB_NO_PDL=1 python tools/slab_kernel_times.py 8192 8 2>&1 | grep -E "rows_inv|spread|sum of"
for w in bearing300 ens1024 flap600; do python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(d['value']), round(d['ms_per_step']*1e3,1), d['stage_us'])"; done
