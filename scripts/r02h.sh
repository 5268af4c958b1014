#!/bin/bash
set -u
OUT=gpurun_out
export NCCL_DEBUG=WARN
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r02h_bench2.log 2>&1; echo "bench rc=$?"; grep -v "^{" $OUT/r02h_bench2.log | tail -60 | cut -c1-300; grep "^{" $OUT/r02h_bench2.log | tail -c 1500
