#!/bin/bash
# round 2, GPU run a: baseline sanity, hit-worker variant (first run ever, every step under its own timeout), ncu of
# the mAP passes and the pack kernel
set -u
OUT=gpurun_out
echo "== baseline gpu tests"
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== default kernel, phases"
timeout 120 python scripts/phase_times.py 2>&1 | tail -12
echo "== WK parity"
CMH_TC_WORKERS=1 timeout 180 python -m pytest tests -x -q -m gpu -k "tc_ or sharded or stripes or finalize or large or config4" 2>&1 | tail -4 || echo "WK parity failed / timed out rc=$?"
echo "== WK phases"
CMH_TC_WORKERS=1 timeout 120 python scripts/phase_times.py 2>&1 | tail -12 || echo "WK phases failed rc=$?"
echo "== map phase timings"
for c in c2-64 c3 c1; do CFG=$c timeout 120 python scripts/map_phase.py 2>&1 | tail -4; done
echo "== ncu map kernels"
for c in c2-64 c3; do
  for k in hist_tile_kernel rank_tile_kernel; do
    CFG=$c timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $OUT/r02a_${c}_$k python scripts/map_phase.py > $OUT/r02a_ncu_${c}_$k.log 2>&1
    echo "capture $c $k rc=$?"
  done
done
CFG=c2-64 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/r02a_map_launches.csv python scripts/map_phase.py > $OUT/r02a_map_launches.log 2>&1
echo "== pack"
timeout 120 python scripts/pack_bench.py 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pack_ -s 6 -c 2 -f -o $OUT/r02a_pack python scripts/pack_bench.py > $OUT/r02a_ncu_pack.log 2>&1
echo "pack capture rc=$?"
echo "== sanitizer (small TC cases)"
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests -x -q -m gpu -k "tc_topk_matches_popc_path and 5000 or tc_edge_cases" > $OUT/r02a_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 $OUT/r02a_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests -x -q -m gpu -k "tc_topk_matches_popc_path and 5000" > $OUT/r02a_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 $OUT/r02a_racecheck.log
ls -la $OUT | tail -12
