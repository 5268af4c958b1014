#!/bin/bash
set -u
OUT=gpurun_out
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 8 --warmup 3 > $OUT/r02i_bench2.log 2>&1; echo "bench rc=$?"; grep -v "^{" $OUT/r02i_bench2.log | tail -5 | cut -c1-300; grep "^{" $OUT/r02i_bench2.log | tail -c 1200
