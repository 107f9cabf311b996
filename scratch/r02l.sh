#!/bin/bash
# round 2, call l: first run of the lane-parallel verifier (pairing VM) on the GPU: parity test, timings, launch durations
set -u
mkdir -p gpurun_out
O=gpurun_out/r02l
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane_parallel or snarkjs or multi_message or batch_proofs_bit_equal" > ${O}_pytest.log 2>&1; echo "verifier tests exit $?" | tee ${O}_summary.txt
tail -15 ${O}_pytest.log
timeout 600 python scratch/verify_timing.py > ${O}_timing.log 2>&1; echo "timing exit $?" | tee -a ${O}_summary.txt
grep -E "ms|program" ${O}_timing.log | tee -a ${O}_summary.txt
tail -5 ${O}_timing.log
