#!/bin/bash
cd "$(dirname "$0")/.."
for sh in "2048 1024" "4096 1024" "8192 64" "512 8192" "600 600" "1024 2048"; do
python tools/sweep_check.py $sh 2>&1 | tail -2 | head -1
JAXIB_COL_SWEEP=0 python tools/sweep_check.py $sh 2>&1 | tail -2 | head -1
done
