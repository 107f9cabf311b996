#!/bin/bash
# round 2, call ab: lazy-reduction Fq2 arithmetic in the G2 accumulate kernel (A/B), reference C examples on the GPU
set -u
mkdir -p gpurun_out
O=gpurun_out/r02ab
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "field_ops or known_answer or bit_equal or production" > ${O}_pytest_sel.log 2>&1; echo "selected tests exit $?" | tee ${O}_summary.txt
tail -3 ${O}_pytest_sel.log | tee -a ${O}_summary.txt
timeout 900 python -m pytest tests/test_abi_exports.py -m gpu -x -q > ${O}_pytest_abi.log 2>&1; echo "abi tests exit $?" | tee -a ${O}_summary.txt
tail -3 ${O}_pytest_abi.log | tee -a ${O}_summary.txt
for v in 1 0; do
  echo "RLN_B200_G2_LAZY=$v" | tee -a ${O}_summary.txt
  RLN_B200_G2_LAZY=$v timeout 300 python scratch/stage_breakdown.py > ${O}_stages_$v.log 2>&1; grep -E "^(1|256|4096) " ${O}_stages_$v.log | tee -a ${O}_summary.txt
done
