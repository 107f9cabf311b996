#!/bin/bash
# round 2, call ac: lazy-reduction Fq2 arithmetic (A/B) x co-resident G1/G2 accumulate (A/B)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02ac
timeout 120 python scratch/lazy_debug.py > ${O}_lazy_debug.log 2>&1; head -10 ${O}_lazy_debug.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "field_ops or known_answer or bit_equal or production" > ${O}_pytest_sel.log 2>&1; echo "selected tests exit $?" | tee ${O}_summary.txt
tail -3 ${O}_pytest_sel.log | tee -a ${O}_summary.txt
for mix in 0 1; do for lazy in 0 1; do
  echo "RLN_B200_MSM_MIX=$mix RLN_B200_G2_LAZY=$lazy" | tee -a ${O}_summary.txt
  RLN_B200_MSM_MIX=$mix RLN_B200_G2_LAZY=$lazy timeout 300 python scratch/stage_breakdown.py > ${O}_stages_${mix}${lazy}.log 2>&1; grep -E "^(256|4096) " ${O}_stages_${mix}${lazy}.log | tee -a ${O}_summary.txt
done; done
