#!/bin/bash
cd "$(dirname "$0")/.."
run() { env "$@" python tools/step_time.py ${WL:-ib2048} 40 2>&1 | tail -1; }
JAXIB_OVERLAP_TRACE=1 JAXIB_IB_OVERLAP=1 python tools/step_time.py ib2048 6 2>&1 | grep trace | tail -8
JAXIB_OVERLAP_TRACE=1 JAXIB_IB_OVERLAP=1 JAXIB_K1_RPC=7 python tools/step_time.py ib2048 6 2>&1 | grep trace | tail -4
run JAXIB_IB_OVERLAP=0 JAXIB_COLS_HALVES=1
run JAXIB_IB_OVERLAP=0 JAXIB_COLS_HALVES=2
run JAXIB_IB_OVERLAP=1 JAXIB_K1_RPC=7
for h in 1 2; do JAXIB_IB_OVERLAP=0 JAXIB_COLS_HALVES=$h python tools/stage_times.py ib2048 30 2>&1 | tail -1; done
