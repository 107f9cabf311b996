#!/bin/bash
# round 2, call v: fused witness program on FOUR slots per bundle (one warp per scheduler) — per-stage times only
set -u
mkdir -p gpurun_out
O=gpurun_out/r02v
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "known_answer or batch_proofs_bit_equal or multi_message" > ${O}_pytest.log 2>&1; echo "subset exit $?" | tee ${O}_summary.txt
tail -2 ${O}_pytest.log
timeout 300 python scratch/stage_breakdown.py > ${O}_stages.log 2>&1; grep -E "^(1|4|32|256|4096) |generate|verify" ${O}_stages.log | tee -a ${O}_summary.txt
