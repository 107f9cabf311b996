#!/bin/bash
# round 2, call p: level records staged through a shared-memory ring by bulk copies, sum routine out of line again
set -u
mkdir -p gpurun_out
O=gpurun_out/r02p
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane_parallel or snarkjs or multi_message" > ${O}_pytest.log 2>&1; echo "verifier tests exit $?" | tee ${O}_summary.txt
tail -3 ${O}_pytest.log
timeout 600 python scratch/verify_timing.py > ${O}_timing.log 2>&1; echo "timing exit $?" | tee -a ${O}_summary.txt
grep -E "ms|program" ${O}_timing.log | tee -a ${O}_summary.txt
timeout 300 python scratch/vm_trace.py > ${O}_trace.txt 2>&1; tail -30 ${O}_trace.txt
