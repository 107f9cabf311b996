#!/bin/bash
# round 2, call w: witness kernel (8 slots, fused program, third operand fetched first): stage times + ncu capture of k_witness with source
set -u
mkdir -p gpurun_out
O=gpurun_out/r02w
timeout 300 python scratch/stage_breakdown.py > ${O}_stages.log 2>&1; grep -E "^(1|4|32|256|4096) |generate|verify" ${O}_stages.log | tee ${O}_summary.txt
RLN_B200_WINDOW_BITS=8 timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_witness$|k_witness\(' -s 2 -c 1 -o ${O}_witness python scratch/single_proof.py > ${O}_ncu.log 2>&1; echo "ncu exit $?" | tee -a ${O}_summary.txt
ls -la ${O}_witness.ncu-rep
