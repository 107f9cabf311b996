#!/bin/bash
# round 2, call ak: --set full capture of the G2 accumulate kernel alone (the multi-kernel capture of call e returned NaN counters for it)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02ak
RLN_BENCH_GLOBAL_BATCH=4096 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_msm_accum' --launch-skip 7 -c 1 \
    -o ${O}_g2 python bench.py --profile --steps 1 --warmup 3 > ${O}_ncu.log 2>&1; echo "ncu exit $?" | tee ${O}_summary.txt
ncu -i ${O}_g2.ncu-rep --page raw --csv > ${O}_g2_raw.csv 2>/dev/null; wc -c ${O}_g2_raw.csv | tee -a ${O}_summary.txt
python scratch/ncu_summarise.py ${O}_g2_raw.csv ${O}_ncu_full_g2.json | tee -a ${O}_summary.txt
