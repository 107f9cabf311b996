#!/bin/bash
# round 2, call n: pairing VM with single-level exponentiations and four-lane sums: parity, timing, per-level trace, ncu capture
set -u
mkdir -p gpurun_out
O=gpurun_out/r02n
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane_parallel or snarkjs or multi_message or batch_proofs_bit_equal" > ${O}_pytest.log 2>&1; echo "verifier tests exit $?" | tee ${O}_summary.txt
tail -5 ${O}_pytest.log
timeout 600 python scratch/verify_timing.py > ${O}_timing.log 2>&1; echo "timing exit $?" | tee -a ${O}_summary.txt
grep -E "ms|program" ${O}_timing.log | tee -a ${O}_summary.txt
timeout 300 python scratch/vm_trace.py > ${O}_trace.txt 2>&1; tail -45 ${O}_trace.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_verify_vm -s 2 -c 1 -o ${O}_vm python scratch/vm_single.py > ${O}_ncu.log 2>&1; echo "ncu exit $?" | tee -a ${O}_summary.txt
ls -la gpurun_out | tail -5
