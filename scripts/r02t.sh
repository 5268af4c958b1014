#!/bin/bash
set -u
OUT=gpurun_out
echo "== any-code-length tests"
timeout 600 python -m pytest tests -x -q -s -m gpu -k "any_code_length or short_codes or tc_" > $OUT/r02t_pytest.log 2>&1; tail -3 $OUT/r02t_pytest.log; grep -E "^(E |bits=)" $OUT/r02t_pytest.log | head -30
echo "== variants by code length"
timeout 900 python scripts/variants_by_bits.py > $OUT/r02t_variants.log 2>&1; echo "rc=$?"; tail -8 $OUT/r02t_variants.log
