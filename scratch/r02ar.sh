#!/bin/bash
# round 2, call ar: four lanes per (proof, group) in the partial-sum reduction of large batches
set -u
mkdir -p gpurun_out
O=gpurun_out/r02ar
: > ${O}_summary.txt
timeout 300 python scratch/stage_breakdown.py > ${O}_stages.log 2>&1; grep -E "^(1|32|256|4096) " ${O}_stages.log | tee -a ${O}_summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "known_answer or bit_equal or production or partial or multi_message" > ${O}_pytest_sel.log 2>&1; echo "selected tests exit $?" | tee -a ${O}_summary.txt
tail -3 ${O}_pytest_sel.log | tee -a ${O}_summary.txt
