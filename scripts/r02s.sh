#!/bin/bash
set -u
OUT=gpurun_out
echo "== gpu tests: f4 first, then the whole suite"
timeout 300 python -m pytest tests -x -q -m gpu -k "set_map" > $OUT/r02s_f4.log 2>&1; tail -3 $OUT/r02s_f4.log; grep -E "^E " $OUT/r02s_f4.log | head -20
timeout 900 python -m pytest tests -q -m gpu > $OUT/r02s_pytest.log 2>&1; tail -3 $OUT/r02s_pytest.log; grep -E "^(E |FAILED)" $OUT/r02s_pytest.log | head -20
