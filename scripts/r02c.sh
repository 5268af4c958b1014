#!/bin/bash
# round 2, GPU run c: full parity suite (C orchestration, lane design, new pack kernels), design comparison of the mAP
# passes, pack GB/s, hit-path A/B, one shard of an 8-GPU search on one GPU
set -u
OUT=gpurun_out
echo "== gpu tests"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "== map designs"
timeout 300 python scripts/map_designs.py 2>&1 | grep -v "^{" | tail -14
echo "== pack"
timeout 120 python scripts/pack_bench.py 2>&1 | tail -2
echo "== workers A/B"
timeout 300 python scripts/workers_ab.py 2>&1 | tail -4
echo "== shard of 8 on one GPU (loopback)"
WORLD=8 timeout 300 python scripts/shard_emul.py 2>&1 | tail -3
WORLD=2 timeout 300 python scripts/shard_emul.py 2>&1 | tail -3
