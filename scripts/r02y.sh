#!/bin/bash
set -u
OUT=gpurun_out
echo "== full gpu suite"
timeout 900 python -m pytest tests -q -m gpu > $OUT/r02y_pytest.log 2>&1; tail -3 $OUT/r02y_pytest.log; grep -E "^(E |FAILED)" $OUT/r02y_pytest.log | head -20
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== default bench"
timeout 900 python bench.py > $OUT/r02y_bench1.log 2>&1; echo "bench rc=$?"; grep -v "^{" $OUT/r02y_bench1.log | tail -5
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r02y_ref.log 2>&1; echo "ref rc=$?"; tail -c 600 $OUT/r02y_ref.log
python - <<'PY'
import json
line=[l for l in open('gpurun_out/r02y_bench1.log') if l.startswith('{')][-1]
d=json.loads(line)
print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches','parity_check')})
print('e2e', {k:v for k,v in d['e2e'].items() if k not in ('note','search_phase_ms')})
print('roofline frac', d['roofline']['frac'], 'clocks', d['clocks'])
for k,v in (d.get('also') or {}).items():
    if isinstance(v, dict): print(k, {kk:vv for kk,vv in v.items() if kk in ('ms','ms_per_call','value','unit','frac','error','items_per_s','seconds')} or list(v)[:8])
PY
