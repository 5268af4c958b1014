#!/bin/bash
set -u
OUT=gpurun_out
timeout 400 python scripts/dbg32c.py 2>&1 | cut -c1-170
timeout 300 python scripts/dbg32b.py 2>&1 | cut -c1-170
echo "== tc tests"
timeout 600 python -m pytest tests -x -q -m gpu -k "tc_ or stream or config4 or stripes or async or benchmarked or any_code or ternary_database or short_codes" > $OUT/r03d_pytest.log 2>&1; tail -3 $OUT/r03d_pytest.log; grep -E "^(E |FAILED)" $OUT/r03d_pytest.log | head
timeout 600 python bench.py --steps 10 --warmup 3 --no-also --no-cpu-baseline > $OUT/r03d_bench.log 2>&1
python - <<'PY'
import json
line=[l for l in open('gpurun_out/r03d_bench.log') if l.startswith('{')][-1]
d=json.loads(line)
print(round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'kernel', round(d['roofline']['kernel_ms_per_step'],2), {k:round(v,2) for k,v in d['phase_ms_per_step'].items()}, d['parity_check']['equal'], d['clocks']['sm_mhz'])
PY
