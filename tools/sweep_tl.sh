#!/bin/bash
cd "$(dirname "$0")/.."
python tools/timeline.py ib2048 12 2>&1 | tail -12
JAXIB_K4_EARLY=1 python tools/timeline.py ib2048 12 2>&1 | tail -12
python tools/timeline.py slab8192 8 2>&1 | tail -12
