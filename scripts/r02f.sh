#!/bin/bash
# round 2, GPU run f (2 GPUs): the library's NCCL transport for real - bench at N=2 incl. the sharded mAP leg
set -u
OUT=gpurun_out
echo "== gpu tests (1 GPU visible to pytest)"
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== bench n=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r02f_bench2.log 2>&1; echo "bench rc=$?"; tail -c 2500 $OUT/r02f_bench2.log
