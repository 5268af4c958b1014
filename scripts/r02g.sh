#!/bin/bash
# round 2, GPU run g (2 GPUs): step-by-step check of the NCCL transport, then a short bench at N=2 - tight timeouts
set -u
OUT=gpurun_out
export NCCL_DEBUG=WARN
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/nccl_smoke.py > $OUT/r02g_smoke.log 2>&1; echo "smoke rc=$?"; grep -E "^\[[01]\]|Error|error|Traceback" $OUT/r02g_smoke.log | tail -30
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-also > $OUT/r02g_bench2.log 2>&1; echo "bench rc=$?"; tail -c 1800 $OUT/r02g_bench2.log
