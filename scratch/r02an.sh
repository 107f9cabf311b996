#!/bin/bash
# round 2, call an: split-bucket combine with 4 / 8 / 16 lanes per bucket at small n (on top of the one-warp Horner pass of call am)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02an
: > ${O}_summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm" > ${O}_pytest_msm.log 2>&1; echo "msm tests exit $?" | tee -a ${O}_summary.txt
tail -2 ${O}_pytest_msm.log | tee -a ${O}_summary.txt
for lg in 16 18 20 22 22 24; do timeout 120 python scratch/msm_profile.py $lg 2>/dev/null | tee -a ${O}_summary.txt; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_msm16_launches.csv python scratch/msm_profile.py 16 > /dev/null 2>&1
grep -E "k_horner|k_bucket_combine" ${O}_msm16_launches.csv | tail -2 | cut -d, -f5,15 | tee -a ${O}_summary.txt
