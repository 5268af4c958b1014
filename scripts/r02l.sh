#!/bin/bash
set -u
OUT=gpurun_out
echo "== gpu tests"
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/r02l_pytest.log 2>&1; tail -4 $OUT/r02l_pytest.log; grep -E "^E " $OUT/r02l_pytest.log | head -12
echo "== shard of 8 (loopback)"
WORLD=8 timeout 200 python scripts/shard_emul.py 2>&1 | tail -2
echo "== bench n=1"
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/r02l_bench1.log 2>&1; echo "bench rc=$?"; grep -v "^{" $OUT/r02l_bench1.log | tail -5; grep "^{" $OUT/r02l_bench1.log | tail -c 300
