#!/bin/bash
# First GPU run of the experimental hit-worker variant of tc_collect_kernel (CMH_TC_WORKERS=1).
# Every step runs under its own timeout: a deadlock in the new queue protocol must not hang the box.
#   gpurun --timeout 600 -- 'bash scripts/gpu_workers_round.sh'
set -u
export CMH_TC_WORKERS=1
echo "== parity (tensor-core search, sharded, stripes) with hit workers"
timeout 120 python -m pytest tests -x -q -m gpu -k "tc_ or sharded or stripes or finalize or large" 2>&1 | tail -4 || echo "parity run failed or timed out (rc=$?)"
echo "== phases, 8192 x 100M"
timeout 90 python scripts/phase_times.py 2>&1 | tail -11 || echo "phase run failed or timed out (rc=$?)"
echo "== hit-path cost by parts"
timeout 120 python scripts/hit_probe.py 2>&1 | tail -8 || echo "probe run failed or timed out (rc=$?)"
unset CMH_TC_WORKERS
echo "== reference: the default kernel, phases"
timeout 90 python scripts/phase_times.py 2>&1 | tail -11
