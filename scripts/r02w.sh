#!/bin/bash
N=${1:-2}
for pc in 0 1; do
  CMH_E2E_DEPTH=1 CMH_E2E_PIECES=$pc timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$pc bench.py --gpus $N --steps 10 --warmup 3 --no-also > gpurun_out/r02w_n${N}_p$pc.log 2>&1
  python - gpurun_out/r02w_n${N}_p$pc.log $pc <<'PY'
import json,sys
line=[l for l in open(sys.argv[1]) if l.startswith('{')][-1]
d=json.loads(line)
print('pieces',sys.argv[2],'N',d['n_gpus'],'value ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),'upload ms',round(d['e2e']['shard_upload_ms'],3), d['parity_check']['equal'])
PY
done
