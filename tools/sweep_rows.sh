#!/bin/bash
# FFT kernel generations (per-stage event timings, L2 flushed)
cd "$(dirname "$0")/.."
run() { env "$@" python tools/stage_times.py ${WL:-ib2048} 30 2>&1 | tail -1; }
run JAXIB_ROWS=fast JAXIB_COLS=fast
run JAXIB_ROWS=fast
run JAXIB_ROWS=team
WL=ib4096 run JAXIB_ROWS=fast
WL=ib4096 run JAXIB_ROWS=team
WL=ib1024 run JAXIB_ROWS=fast
WL=ib1024 run JAXIB_ROWS=team
WL=slab8192 run JAXIB_ROWS=fast
WL=slab8192 run JAXIB_ROWS=team
