#!/bin/bash
set -u
OUT=gpurun_out
echo "== tc tests (32-bit operand variant)"
timeout 600 python -m pytest tests -x -q -m gpu -k "tc_ or stream or config4 or stripes or async or benchmarked or any_code or ternary_database or short_codes" > $OUT/r03c_pytest.log 2>&1; tail -3 $OUT/r03c_pytest.log; grep -E "^(E |FAILED)" $OUT/r03c_pytest.log | head
echo "== variants by code length"
Q=8192 D=50000000 REPS=3 timeout 900 python scripts/variants_by_bits.py > $OUT/r03c_variants.log 2>&1; echo "rc=$?"; grep -E "^(16|32|48|64|96|128) " $OUT/r03c_variants.log
