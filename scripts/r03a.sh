#!/bin/bash
for r in 1 2 3; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-also --no-cpu-baseline > gpurun_out/r03a_$r.log 2>&1
  python - gpurun_out/r03a_$r.log <<'PY'
import json,sys
line=[l for l in open(sys.argv[1]) if l.startswith('{')][-1]
d=json.loads(line)
print(round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['ms_each_step'], d['clocks']['sm_min_mhz'], d['clocks']['sm_mhz'], d['clocks']['power_w_max'])
PY
done
