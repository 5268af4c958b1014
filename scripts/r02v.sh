#!/bin/bash
# e2e pipeline depth A/B at N ranks: prints value / e2e / upload ms per depth
N=${1:-2}
for d in 1 2; do
  if [ "$N" = "1" ]; then
    CMH_E2E_DEPTH=$d timeout 600 python bench.py --steps 10 --warmup 3 --no-also > gpurun_out/r02v_n${N}_d$d.log 2>&1
  else
    CMH_E2E_DEPTH=$d timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$d bench.py --gpus $N --steps 10 --warmup 3 --no-also > gpurun_out/r02v_n${N}_d$d.log 2>&1
  fi
  python - gpurun_out/r02v_n${N}_d$d.log $d <<'PY'
import json,sys
line=[l for l in open(sys.argv[1]) if l.startswith('{')][-1]
d=json.loads(line)
print('depth',sys.argv[2],'N',d['n_gpus'],'value ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),'upload ms',round(d['e2e']['shard_upload_ms'],3), d['parity_check']['equal'])
PY
done
