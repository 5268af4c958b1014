#!/bin/bash
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 --no-also > gpurun_out/r02x_n${N}.log 2>&1
python - gpurun_out/r02x_n${N}.log <<'PY'
import json,sys
line=[l for l in open(sys.argv[1]) if l.startswith('{')][-1]
d=json.loads(line)
e=d['e2e']
print(e.get("pinned_host_numa")); print("N",d["n_gpus"],"value ms",round(d['ms_per_step'],3),'e2e ms',round(e['ms_per_step'],3),'upload',round(e['shard_upload_ms'],3),'search',round(e['search_ms_per_step'],3),'span',round(e['first_search_start_to_last_search_end_ms_per_step'],3), d['phase_ms_per_step'])
PY
