#!/bin/bash
cd "$(dirname "$0")/.."
for w in ib2048 slab8192; do python tools/stage_times.py $w 20 2>&1 | tail -1; python tools/step_time.py $w 20 2>&1 | tail -1; done
