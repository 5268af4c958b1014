#!/bin/bash
# round 2 profiles: launch list of one bench step, full captures of the dominant kernels (after the plain runs exited 0)
set -u
OUT=gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-also"
$CMD > $OUT/r02p_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/r02p_plain.log; exit 1; }
tail -c 200 $OUT/r02p_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/r02p_launches.csv $CMD > $OUT/r02p_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_collect_kernel -s 23 -c 1 -f -o $OUT/r02p_tc_collect_kernel $CMD > $OUT/r02p_ncu_tc.log 2>&1; echo "tc_collect capture rc=$?"
for c in c2-64 c3; do
  for k in hist_lane_kernel rank_lane_kernel; do
    CFG=$c timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $OUT/r02p_${c}_$k python scripts/map_phase.py > $OUT/r02p_ncu_${c}_$k.log 2>&1; echo "capture $c $k rc=$?"
  done
done
CFG=c2-64 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/r02p_map_launches.csv python scripts/map_phase.py > $OUT/r02p_map_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pack_ -s 6 -c 2 -f -o $OUT/r02p_pack python scripts/pack_bench.py > $OUT/r02p_ncu_pack.log 2>&1; echo "pack capture rc=$?"
ls -la $OUT | grep r02p
