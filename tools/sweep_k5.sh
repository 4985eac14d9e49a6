#!/bin/bash
# K5 variants (per-stage event timings, L2 flushed)
cd "$(dirname "$0")/.."
run() { env "$@" python tools/stage_times.py ${WL:-ib2048} 30 2>&1 | tail -1; }
run JAXIB_COLS=fast
run JAXIB_COLS=team
run JAXIB_COLS_CPC=8
run JAXIB_COLS_CPC=4
WL=ib4096 run JAXIB_COLS=fast
WL=ib4096 run JAXIB_COLS=team
WL=ib1024 run JAXIB_COLS=fast
WL=ib1024 run JAXIB_COLS=team
