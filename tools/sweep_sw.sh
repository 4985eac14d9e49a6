#!/bin/bash
cd "$(dirname "$0")/.."
run() { env "$@" python tools/step_time.py ${WL:-ib2048} 40 2>&1 | tail -1; }
st() { env "$@" python tools/stage_times.py ${WL:-ib2048} 30 2>&1 | tail -1; }
st JAXIB_X=1
run JAXIB_X=1
run JAXIB_SWEEP_CS=4
st JAXIB_SWEEP_CS=4
run JAXIB_SWEEP_CS=2
WL=ib4096 st JAXIB_X=1
WL=ib4096 run JAXIB_X=1
WL=ib1024 run JAXIB_X=1
WL=ib1024 run JAXIB_SWEEP_CS=8
WL=slab8192 st JAXIB_X=1
WL=slab8192 run JAXIB_X=1
WL=flap600 st JAXIB_X=1
WL=ens1024 st JAXIB_X=1
WL=channel2048 st JAXIB_X=1
