#!/bin/bash
# round 2, call z: folded assembly (no variable-base multiplication for ≤ 32 full proofs) + proof values read off the witness:
# full GPU suite, per-stage times with the fold on and off
set -u
mkdir -p gpurun_out
O=gpurun_out/r02z
timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "suite exit $?" | tee ${O}_summary.txt
tail -3 ${O}_pytest.log
for v in 1 0; do
  echo "RLN_B200_ASSEMBLE_FOLD=$v" | tee -a ${O}_summary.txt
  RLN_B200_ASSEMBLE_FOLD=$v timeout 300 python scratch/stage_breakdown.py > ${O}_stages_$v.log 2>&1; grep -E "^(1|4|32|256) |generate|verify" ${O}_stages_$v.log | tee -a ${O}_summary.txt
done
