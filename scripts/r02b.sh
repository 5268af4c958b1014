#!/bin/bash
# round 2, GPU run b: the C-owned orchestration (cmh_topk_tc) - full parity suite, phases, a short bench
set -u
OUT=gpurun_out
echo "== gpu tests"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "== phases (hybrid: pilot = draining warps, main = hit workers)"
timeout 180 python scripts/phase_times.py 2>&1 | tail -5
echo "== phases, all launches on the draining-warp kernel"
CMH_TC_WORKERS=0 timeout 180 python scripts/phase_times.py 2>&1 | tail -4
echo "== phases, sample 16K rows"
SAMPLE_ROWS=16384 timeout 180 python scripts/phase_times.py 2>&1 | tail -4
echo "== bench n=1"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/r02b_bench1.log 2>&1; echo "bench rc=$?"; tail -c 1500 $OUT/r02b_bench1.log
