#!/bin/bash
# round 2, GPU run d: full parity suite, CTA waves per tc_collect launch, ncu of the lane kernels, pack, bench
set -u
OUT=gpurun_out
echo "== gpu tests"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== waves per launch: N=1 full search and one shard of 8"
for w in 1 2 4; do
  echo "-- CMH_TC_WAVES=$w"
  CMH_TC_WAVES=$w timeout 200 python scripts/phase_times.py 2>&1 | tail -3
  CMH_TC_WAVES=$w WORLD=8 timeout 200 python scripts/shard_emul.py 2>&1 | tail -1
done
echo "== pack"
timeout 120 python scripts/pack_bench.py 2>&1 | tail -2
echo "== map designs"
timeout 300 python scripts/map_designs.py 2>&1 | grep -v "^{" | tail -12
echo "== ncu lane kernels"
for c in c2-64 c3; do
  for k in hist_lane_kernel rank_lane_kernel; do
    CFG=$c timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $OUT/r02d_${c}_$k python scripts/map_phase.py > $OUT/r02d_ncu_${c}_$k.log 2>&1
    echo "capture $c $k rc=$?"
  done
done
echo "== bench n=1"
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/r02d_bench1.log 2>&1; echo "bench rc=$?"; tail -c 600 $OUT/r02d_bench1.log
