#!/bin/bash
set -u
OUT=gpurun_out
echo "== tc tests (threshold in the drain)"
timeout 600 python -m pytest tests -x -q -m gpu -k "tc_ or stream or config4 or stripes or async or benchmarked or any_code or ternary_database or short_codes" > $OUT/r03b_pytest.log 2>&1; tail -3 $OUT/r03b_pytest.log; grep -E "^(E |FAILED)" $OUT/r03b_pytest.log | head
for mode in drain mma drain mma; do
  CMH_TC_BIAS=$mode timeout 600 python bench.py --steps 10 --warmup 3 --no-also --no-cpu-baseline > $OUT/r03b_$mode.log 2>&1
  python - $OUT/r03b_$mode.log $mode <<'PY'
import json,sys
line=[l for l in open(sys.argv[1]) if l.startswith('{')][-1]
d=json.loads(line)
print(sys.argv[2], round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'kernel', round(d['roofline']['kernel_ms_per_step'],2), {k:round(v,2) for k,v in d['phase_ms_per_step'].items()}, {k:round(v,2) for k,v in d['roofline']['in_situ_ceilings_ms'].items()}, d['parity_check']['equal'], d['parity_check']['n_fail'], d['clocks']['sm_mhz'])
PY
done
