#!/bin/bash
# round 2, call at: ncu launch list of one device batch of the final code (the same command as the bench, one step)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02at
RLN_BENCH_GLOBAL_BATCH=4096 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file ${O}_launches.csv \
    python bench.py --profile --steps 1 --warmup 3 > ${O}_ncu_list.log 2>&1; echo "ncu launch list exit $?" | tee ${O}_summary.txt
wc -l ${O}_launches.csv | tee -a ${O}_summary.txt
