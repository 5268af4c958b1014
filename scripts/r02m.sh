#!/bin/bash
set -u
OUT=gpurun_out
echo "== gpu tests"
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/r02m_pytest.log 2>&1; tail -4 $OUT/r02m_pytest.log; grep -E "^E " $OUT/r02m_pytest.log | head -12
echo "== bench n=1"
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/r02m_bench1.log 2>&1; echo "bench rc=$?"; grep -v "^{" $OUT/r02m_bench1.log | tail -5; grep "^{" $OUT/r02m_bench1.log | tail -c 300
