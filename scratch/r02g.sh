#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02g
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "suite exit $?" | tee ${O}_summary.txt
tail -4 ${O}_pytest.log
for lg in 16 18 20 22 24; do timeout 120 python scratch/msm_profile.py $lg 2>/dev/null | tee -a ${O}_summary.txt; done
for v in 0 1; do
  RLN_B200_NTT_TILED=$v timeout 300 python scratch/stage_breakdown.py > ${O}_stages_ntt$v.log 2>&1; echo "NTT_TILED=$v" | tee -a ${O}_summary.txt
  grep -E "^(1|256|4096) " ${O}_stages_ntt$v.log | tee -a ${O}_summary.txt
done
timeout 600 python scratch/affine_probe.py 2>&1 | tee -a ${O}_summary.txt
