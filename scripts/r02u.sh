#!/bin/bash
set -u
OUT=gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --steps 10 --warmup 3 --no-also > $OUT/r02u_bench1.log 2>&1; echo "bench rc=$?"; grep -v "^{" $OUT/r02u_bench1.log | tail -5
  F=$OUT/r02u_bench1.log
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-also > $OUT/r02u_bench$N.log 2>&1; echo "bench rc=$?"; grep -v "^{" $OUT/r02u_bench$N.log | tail -5
  F=$OUT/r02u_bench$N.log
fi
python - $F <<'PY'
import json,sys
line=[l for l in open(sys.argv[1]) if l.startswith('{')][-1]
d=json.loads(line)
print({k:d[k] for k in ('n_gpus','value','ms_per_step','parity_check')}, 'e2e', d['e2e']['ms_per_step'], d['clocks'], d['phase_ms_per_step'])
PY
