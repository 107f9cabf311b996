#!/bin/bash
# round 2, call af: G2 accumulate — both coordinates of every Fq2 product side by side (four carry chains per warp), 2 / 3 CTAs per SM
set -u
mkdir -p gpurun_out
O=gpurun_out/r02af
: > ${O}_summary.txt
for cfg in "0 2" "1 2" "0 3" "1 3"; do set -- $cfg
  echo "RLN_B200_G2_PAIR=$1 RLN_B200_G2_BLOCKS=$2" | tee -a ${O}_summary.txt
  RLN_B200_G2_PAIR=$1 RLN_B200_G2_BLOCKS=$2 timeout 300 python scratch/stage_breakdown.py > ${O}_stages_$1$2.log 2>&1; grep -E "^(256|4096) " ${O}_stages_$1$2.log | tee -a ${O}_summary.txt
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "field_ops or known_answer or bit_equal or production" > ${O}_pytest_sel.log 2>&1; echo "selected tests exit $?" | tee -a ${O}_summary.txt
tail -3 ${O}_pytest_sel.log | tee -a ${O}_summary.txt
