#!/bin/bash
cd "$(dirname "$0")/.."
run() { env "$@" python tools/step_time.py ${WL:-ib2048} 40 2>&1 | tail -1; }
run JAXIB_K4_EARLY=0
run JAXIB_K4_EARLY=1
run JAXIB_K4_EARLY=1 JAXIB_K4_EARLY_THREADS=1024
run JAXIB_K4_EARLY=1 JAXIB_K4_EARLY_THREADS=256
WL=ib1024 run JAXIB_K4_EARLY=0
WL=ib1024 run JAXIB_K4_EARLY=1
