#!/bin/bash
# round 2, call ah: one-level Karatsuba + fused-shift reduction as the Fq product of the fixed-base MSM unit (A/B against the
# interleaved-carry product; G1 accumulate at 4 and at 3 CTAs per SM)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02ah
: > ${O}_summary.txt
KARA=$PWD/zerokit_b200/lib/librln_b200_kara.so
for cfg in "plain 4" "kara 4" "kara 3" "plain 3"; do set -- $cfg
  echo "lib=$1 RLN_B200_G1_BLOCKS=$2" | tee -a ${O}_summary.txt
  if [ $1 = kara ]; then export RLN_B200_LIB=$KARA; else unset RLN_B200_LIB; fi
  RLN_B200_G1_BLOCKS=$2 timeout 300 python scratch/stage_breakdown.py > ${O}_stages_$1_$2.log 2>&1; grep -E "^(1|256|4096) " ${O}_stages_$1_$2.log | tee -a ${O}_summary.txt
done
export RLN_B200_LIB=$KARA
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "known_answer or bit_equal or production or partial" > ${O}_pytest_sel.log 2>&1; echo "selected tests (kara) exit $?" | tee -a ${O}_summary.txt
tail -3 ${O}_pytest_sel.log | tee -a ${O}_summary.txt
