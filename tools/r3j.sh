#!/bin/bash
cd "$(dirname "$0")/.."
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_slab.py tests/test_gpu_round2.py -m gpu -q -x -k "pressure_solve or column_sweep or full_step or channel or far_field or ensemble or golden" 2>&1 | tail -3
for t in 1 0; do JAXIB_SWEEP_TMA=$t python tools/step_time.py ib2048 40 | tail -1; done
for t in 1 0; do JAXIB_SWEEP_TMA=$t python tools/stage_times.py ib2048 30 | tail -1; done
for t in 1 0; do JAXIB_SWEEP_TMA=$t python tools/step_time.py ib4096 20 | tail -1; done
for t in 1 0; do JAXIB_SWEEP_TMA=$t python tools/step_time.py slab8192 20 | tail -1; done
for t in 1 0; do JAXIB_SWEEP_TMA=$t JAXIB_NO_PDL=1 python tools/slab_kernel_times.py 8192 8 2>&1 | grep -E "cols_sweep|sum of"; done
