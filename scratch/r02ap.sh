#!/bin/bash
# round 2, call ap: compute-sanitizer memcheck / racecheck over the changed kernels of the variable-base MSM tail
set -u
mkdir -p gpurun_out
O=gpurun_out/r02ap
timeout 120 python scratch/sanitize_msm.py > ${O}_plain.log 2>&1; echo "plain exit $?" | tee ${O}_summary.txt
timeout 900 compute-sanitizer --tool memcheck python scratch/sanitize_msm.py > ${O}_memcheck.log 2>&1; echo "memcheck exit $?" | tee -a ${O}_summary.txt
tail -3 ${O}_memcheck.log | tee -a ${O}_summary.txt
timeout 900 compute-sanitizer --tool racecheck python scratch/sanitize_msm.py > ${O}_racecheck.log 2>&1; echo "racecheck exit $?" | tee -a ${O}_summary.txt
tail -3 ${O}_racecheck.log | tee -a ${O}_summary.txt
