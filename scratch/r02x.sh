#!/bin/bash
# round 2, call x: four-warp proof assembly for small batches: full GPU suite + per-stage times
set -u
mkdir -p gpurun_out
O=gpurun_out/r02x
timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "suite exit $?" | tee ${O}_summary.txt
tail -3 ${O}_pytest.log
timeout 300 python scratch/stage_breakdown.py > ${O}_stages.log 2>&1; grep -E "^(1|4|32|256|4096) |generate|verify" ${O}_stages.log | tee -a ${O}_summary.txt
