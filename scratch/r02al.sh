#!/bin/bash
# round 2, call al: launch list of the variable-base MSM at 2^16 and 2^20 on the final code (where does the small-n time go?)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02al
: > ${O}_summary.txt
for lg in 16 20; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_msm${lg}_launches.csv python scratch/msm_profile.py $lg > /dev/null 2>&1; echo "ncu msm$lg exit $?" | tee -a ${O}_summary.txt
done
