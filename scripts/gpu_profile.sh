#!/bin/bash
# Run on the GPU box (via gpurun): launch list + one full ncu capture of each heavy kernel of the bench step.
# Usage: scripts/gpu_profile.sh <tag> [kernel-regex ...]
set -u
TAG=${1:-r01}; shift || true
KERNELS=("$@"); [ ${#KERNELS[@]} -eq 0 ] && KERNELS=(hist_tile_kernel select_tile_kernel)
CMD="python bench.py --steps 1 --warmup 3 --queries 2048 --db-rows 25000000 --no-cpu-baseline --no-also"
OUT=gpurun_out
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
for k in "${KERNELS[@]}"; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $OUT/${TAG}_$k $CMD > $OUT/${TAG}_ncu_$k.log 2>&1
  echo "full capture $k rc=$?"
done
ls -la $OUT
