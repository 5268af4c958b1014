#!/bin/bash
# Run on the GPU box (via gpurun): launch list + one full ncu capture of the dominant kernel of the bench step.
# Usage: scripts/gpu_profile.sh <tag> [kernel-regex] [skip]
set -u
TAG=${1:-r01}; KERNEL=${2:-tc_collect_kernel}; SKIP=${3:-7}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-also"
OUT=gpurun_out
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/${TAG}_plain.log; exit 1; }
tail -c 600 $OUT/${TAG}_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
# main launch of the timed step: 2 tc_collect launches per step, 3 warm-up steps -> skip 7
ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s $SKIP -c 1 -f -o $OUT/${TAG}_$KERNEL $CMD > $OUT/${TAG}_ncu_$KERNEL.log 2>&1
echo "full capture $KERNEL rc=$?"
ls -la $OUT | tail -8
