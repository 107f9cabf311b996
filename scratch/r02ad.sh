#!/bin/bash
# round 2, call ad: co-resident G1/G2 accumulate with matching shared-memory carve-outs
set -u
mkdir -p gpurun_out
O=gpurun_out/r02ad
: > ${O}_summary.txt
for cfg in "0 0 58" "1 0 58" "1 1 58" "1 0 100" "1 1 100"; do set -- $cfg
  echo "RLN_B200_MSM_MIX=$1 RLN_B200_G2_LAZY=$2 RLN_B200_MIX_CARVE=$3" | tee -a ${O}_summary.txt
  RLN_B200_MSM_MIX=$1 RLN_B200_G2_LAZY=$2 RLN_B200_MIX_CARVE=$3 timeout 300 python scratch/stage_breakdown.py > ${O}_stages_$1$2_$3.log 2>&1; grep -E "^(4096) " ${O}_stages_$1$2_$3.log | tee -a ${O}_summary.txt
done
