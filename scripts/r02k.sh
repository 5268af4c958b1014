#!/bin/bash
set -u
echo "== gpu tests"
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== prefix cuts, one GPU"
MODE=one timeout 400 python scripts/sweep_cuts.py 2>&1 | tail -7
echo "== prefix cuts, shard of 8"
MODE=shard WORLD=8 timeout 400 python scripts/sweep_cuts.py 2>&1 | tail -7
