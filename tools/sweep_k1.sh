#!/bin/bash
# K1 variants at ib2048 (per-stage event timings, L2 flushed): old marching kernel vs the bulk-copy ring
cd "$(dirname "$0")/.."
run() { env "$@" python tools/stage_times.py ib2048 30 2>&1 | tail -1; }
run JAXIB_K1=march
run JAXIB_K1=ring
run JAXIB_K1_W=2048
run JAXIB_K1_W=512
run JAXIB_K1_DEPTH=3
run JAXIB_K1_DEPTH=6
run JAXIB_K1_RPC=7
run JAXIB_K1_RPC=28
run JAXIB_K1_W=512 JAXIB_K1_DEPTH=6
