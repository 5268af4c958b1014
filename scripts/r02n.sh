#!/bin/bash
set -u
OUT=gpurun_out
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/r02n_bench8.log 2>&1; echo "bench rc=$?"; grep -v "^{" $OUT/r02n_bench8.log | tail -5 | cut -c1-300; grep "^{" $OUT/r02n_bench8.log | tail -c 300
