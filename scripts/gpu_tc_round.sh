#!/bin/bash
# tensor-core path: parity tests, then the piecewise timing with in-situ probes (run on the GPU box via gpurun)
set -u
TAG=${1:-r01d}
OUT=gpurun_out
python -m pytest tests -x -q -m gpu -k "tc_" > $OUT/${TAG}_pytest_tc.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest_tc.log
python scripts/bench_tc.py > $OUT/${TAG}_tc_full.log 2>&1 || { echo "bench_tc failed"; tail -20 $OUT/${TAG}_tc_full.log; }
tail -1 $OUT/${TAG}_tc_full.log
