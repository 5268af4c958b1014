#!/bin/bash
set -u
OUT=gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "pinned or pack_codes or four_direction or map_k_matches" > $OUT/r02z_pytest.log 2>&1; tail -3 $OUT/r02z_pytest.log; grep -E "^(E |FAILED)" $OUT/r02z_pytest.log | head
timeout 600 python bench.py --steps 5 --warmup 3 --no-also > $OUT/r02z_bench1.log 2>&1; echo "bench rc=$?"; grep -v "^{" $OUT/r02z_bench1.log | tail -5
python - <<'PY'
import json
line=[l for l in open('gpurun_out/r02z_bench1.log') if l.startswith('{')][-1]
d=json.loads(line)
print({k:d[k] for k in ('value','ms_per_step','steps')}, 'e2e', {k:v for k,v in d['e2e'].items() if k not in ('note','search_phase_ms','unit')})
PY
