#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02f
timeout 900 python -m pytest tests -m gpu -x -q -k "msm or production or batch_proofs or known_answer" > ${O}_pytest.log 2>&1; echo "msm+proof tests exit $?" | tee ${O}_summary.txt
tail -5 ${O}_pytest.log
for lg in 16 18 20 22 24; do timeout 120 python scratch/msm_profile.py $lg 2>/dev/null | tee -a ${O}_summary.txt; done
for v in 0 1; do
  RLN_B200_G2_SACC=$v timeout 300 python scratch/stage_breakdown.py > ${O}_stages_sacc$v.log 2>&1; echo "G2_SACC=$v" | tee -a ${O}_summary.txt
  grep -E "^(256|4096) " ${O}_stages_sacc$v.log | tee -a ${O}_summary.txt
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_msm16_launches.csv python scratch/msm_profile.py 16 > /dev/null 2>&1; echo "ncu msm16 exit $?" | tee -a ${O}_summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_msm22_launches.csv python scratch/msm_profile.py 22 > /dev/null 2>&1; echo "ncu msm22 exit $?" | tee -a ${O}_summary.txt
