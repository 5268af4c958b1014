#!/bin/bash
# Run on the GPU box (via gpurun): launch list + full ncu captures of kernels of the bench step.
# Usage: scripts/gpu_profile.sh <tag> [kernel-regex:skip ...]
#   tc_collect_kernel is launched 7 times per step (2 pilot stages + 5 spans); after 3 warm-up steps, skip 23 = the first main span
#   of the timed step, skip 3 = the 4th topk_finalize_kernel (the timed step's)
set -u
TAG=${1:-r01}; shift
PAIRS=${@:-tc_collect_kernel:23 topk_finalize_kernel:3}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-also"
OUT=gpurun_out
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/${TAG}_plain.log; exit 1; }
tail -c 300 $OUT/${TAG}_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
for pair in $PAIRS; do
    KERNEL=${pair%%:*}; SKIP=${pair##*:}
    ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s $SKIP -c 1 -f -o $OUT/${TAG}_$KERNEL $CMD > $OUT/${TAG}_ncu_$KERNEL.log 2>&1
    echo "full capture $KERNEL (skip $SKIP) rc=$?"
done
ls -la $OUT | tail -8
