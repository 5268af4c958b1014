#!/bin/bash
# ncu capture of the tensor-core collect kernel on a reduced workload (run on the GPU box via gpurun)
set -u
TAG=${1:-r01b}; SKIP=${2:-5}
export Q=2048 D=25000000
CMD="python scripts/bench_tc.py"
OUT=gpurun_out
$CMD > $OUT/${TAG}_tc_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/${TAG}_tc_plain.log; exit 1; }
tail -1 $OUT/${TAG}_tc_plain.log
# launch index: skip the timed() warm-up + static-threshold launches; take one launch of the tightened collect
ncu --set full --clock-control none --import-source on -k regex:tc_collect_kernel -s $SKIP -c 1 -f -o $OUT/${TAG}_tc_collect $CMD > $OUT/${TAG}_ncu_tc.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/${TAG}_ncu_tc.log
