#!/bin/bash
cd "$(dirname "$0")/.."
run() { env "$@" python tools/step_time.py ${WL:-ib2048} 40 2>&1 | tail -1; }
run JAXIB_IB_OVERLAP=0
run JAXIB_IB_OVERLAP=1
run JAXIB_IB_OVERLAP=1 JAXIB_K1_W=2048
run JAXIB_IB_OVERLAP=1 JAXIB_K1_W=512
WL=slab8192 run JAXIB_IB_OVERLAP=0
WL=slab8192 run JAXIB_IB_OVERLAP=1
WL=ib4096 run JAXIB_IB_OVERLAP=0
WL=ib4096 run JAXIB_IB_OVERLAP=1
