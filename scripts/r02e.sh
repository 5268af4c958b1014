#!/bin/bash
set -u
OUT=gpurun_out
echo "== gpu tests"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== lane modes (bit0: peers by OR, bit1: 4 sub-histograms)"
for m in 0 1 2 3; do echo "-- CMH_LANE_MODE=$m"; CMH_LANE_MODE=$m DESIGNS=2 CFGS=c1,c2-64,c3,c5 timeout 200 python scripts/map_designs.py 2>&1 | grep -v "^{" | tail -4; done
echo "== tile for reference"
DESIGNS=0 CFGS=c2-16,c2-32 timeout 200 python scripts/map_designs.py 2>&1 | grep -v "^{" | tail -2
echo "== map phase (whole calls)"
for c in c1 c2-64 c3; do CFG=$c timeout 120 python scripts/map_phase.py 2>&1 | tail -4 | head -2; done
