#!/bin/bash
# round 2, call aa: G2 sums beside the G1 sums for small batches, fold up to 16 proofs
# full GPU suite, per-stage times with the fold on and off
set -u
mkdir -p gpurun_out
O=gpurun_out/r02aa
timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "suite exit $?" | tee ${O}_summary.txt
tail -3 ${O}_pytest.log
for v in 1 0; do
  echo "RLN_B200_ASSEMBLE_FOLD=$v" | tee -a ${O}_summary.txt
  RLN_B200_ASSEMBLE_FOLD=$v timeout 300 python scratch/stage_breakdown.py > ${O}_stages_$v.log 2>&1; grep -E "^(1|4|32|256) |generate|verify" ${O}_stages_$v.log | tee -a ${O}_summary.txt
done
