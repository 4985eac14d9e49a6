#!/bin/bash
cd "$(dirname "$0")/.."
ncu --set full --clock-control none --import-source on -k regex:"cols_sweep" -s 5 -c 2 -o gpurun_out/prof_r2v python tools/stage_times.py ib2048 3 > gpurun_out/r2v_ncu.log 2>&1
tail -2 gpurun_out/r2v_ncu.log
