#!/bin/bash
# final validation of round 2 + refreshed launch list / capture of the dominant kernel
set -u
OUT=gpurun_out
echo "== full gpu suite"
timeout 900 python -m pytest tests -q -m gpu > $OUT/r03g_pytest.log 2>&1; tail -2 $OUT/r03g_pytest.log; grep -E "^(E |FAILED)" $OUT/r03g_pytest.log | head -20
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== default-flag bench (with also)"
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/r03g_bench1.log 2>&1; echo "bench rc=$?"; grep -v "^{" $OUT/r03g_bench1.log | tail -3
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-also"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/r03g_launches.csv $CMD > $OUT/r03g_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_collect_kernel -s 23 -c 1 -f -o $OUT/r03g_tc_collect_kernel $CMD > $OUT/r03g_ncu_tc.log 2>&1; echo "tc_collect capture rc=$?"
python - <<'PY'
import json
line=[l for l in open('gpurun_out/r03g_bench1.log') if l.startswith('{')][-1]
d=json.loads(line)
print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches')}, d['parity_check']['equal'], d['parity_check']['n_fail'])
print('e2e', {k:v for k,v in d['e2e'].items() if k not in ('note','search_phase_ms','unit','pinned_host_numa')})
print('roofline frac', d['roofline']['frac'], 'kernel ms', d['roofline']['kernel_ms_per_step'], 'clocks', d['clocks'])
for k,v in (d.get('also') or {}).items():
    if isinstance(v, dict): print(k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms_per_call_i2t','ms_per_step','items_per_s','seconds','error')})
PY
