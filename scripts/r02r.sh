#!/bin/bash
set -u
OUT=gpurun_out
echo "== gpu tests (tc / index subset)"
timeout 600 python -m pytest tests -x -q -m gpu -k "tc_ or stream or config4 or stripes or async or benchmarked" > $OUT/r02r_pytest.log 2>&1; tail -3 $OUT/r02r_pytest.log; grep -E "^E " $OUT/r02r_pytest.log | head
echo "== bench n=1 (no also)"
timeout 600 python bench.py --steps 10 --warmup 3 --no-also > $OUT/r02r_bench1.log 2>&1; echo "bench rc=$?"; grep -v "^{" $OUT/r02r_bench1.log | tail -5
python - <<'PY'
import json
line=[l for l in open('gpurun_out/r02r_bench1.log') if l.startswith('{')][-1]
d=json.loads(line)
print({k:d[k] for k in ('value','ms_per_step','parity_check')}, 'e2e', d['e2e']['ms_per_step'], d['clocks'])
PY
