#!/bin/bash
# round 2, call r: witness program re-associated by readiness (depth 10 004 -> 6 300 bundles), eight slots per bundle: GPU suite + per-stage times, A/B
set -u
mkdir -p gpurun_out
O=gpurun_out/r02r
timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "suite exit $?" | tee ${O}_summary.txt
tail -3 ${O}_pytest.log
for v in 1 0; do
  echo "RLN_B200_WITNESS_REASSOC=$v" | tee -a ${O}_summary.txt
  RLN_B200_WITNESS_REASSOC=$v timeout 300 python scratch/stage_breakdown.py > ${O}_stages_$v.log 2>&1; grep -E "^(1|4|32|256|4096) |generate|verify" ${O}_stages_$v.log | tee -a ${O}_summary.txt
done
