#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests -q -m gpu > $OUT/r03h_pytest.log 2>&1; tail -2 $OUT/r03h_pytest.log; grep -E "^(E |FAILED)" $OUT/r03h_pytest.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/r03h_bench1.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
line=[l for l in open('gpurun_out/r03h_bench1.log') if l.startswith('{')][-1]
d=json.loads(line)
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['parity_check']['equal'], d['parity_check']['n_fail'], 'e2e', round(d['e2e']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])
print({k:(v.get('ms_per_call_i2t') or v.get('ms_per_step') or v.get('items_per_s')) for k,v in (d.get('also') or {}).items() if isinstance(v,dict)})
PY
