#!/bin/bash
set -u
OUT=gpurun_out
echo "== gpu tests"
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/r02q_pytest.log 2>&1; tail -4 $OUT/r02q_pytest.log; grep -E "^E " $OUT/r02q_pytest.log | head -12
echo "== map designs"
DESIGNS=2 CFGS=c1,c2-64,c3,c5 timeout 200 python scripts/map_designs.py 2>&1 | grep -v "^{" | tail -4
echo "== map phase (whole calls)"
for c in c1 c2-64 c3; do CFG=$c timeout 120 python scripts/map_phase.py 2>&1 | tail -4 | head -2; done
echo "== pack"
timeout 120 python scripts/pack_bench.py 2>&1 | tail -1
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
