#!/bin/bash
set -u
OUT=gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "dense_hits or any_code or short_codes" > $OUT/r03e_pytest.log 2>&1; tail -3 $OUT/r03e_pytest.log; grep -E "^(E |FAILED)" $OUT/r03e_pytest.log | head
echo "== variants by code length"
Q=8192 D=20000000 REPS=3 timeout 900 python scripts/variants_by_bits.py > $OUT/r03e_variants.log 2>&1; echo "rc=$?"; grep -E "^(16|32|48|64|96|128) " $OUT/r03e_variants.log | cut -c1-150
