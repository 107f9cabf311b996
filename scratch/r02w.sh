#!/bin/bash
# round 2, call w: ncu capture of k_witness (single proof, 8 slots, fused program) with source-level stall samples
set -u
mkdir -p gpurun_out
O=gpurun_out/r02w
RLN_B200_WINDOW_BITS=8 timeout 600 ncu --set full --import-source on --clock-control none -k k_witness -s 1 -c 1 -o ${O}_witness python scratch/single_proof.py > ${O}_ncu.log 2>&1; echo "ncu exit $?" | tee ${O}_summary.txt
tail -3 ${O}_ncu.log
ls -la ${O}_witness.ncu-rep
