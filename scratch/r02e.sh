#!/bin/bash
# 1-GPU call: suite (with the new verifier), stage breakdown, ncu launch list of one bench step + full captures of the top kernels
set -u
mkdir -p gpurun_out
O=gpurun_out/r02e
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > ${O}_pytest.log 2>&1; echo "suite exit $?" | tee ${O}_summary.txt
tail -15 ${O}_pytest.log
timeout 300 python scratch/stage_breakdown.py > ${O}_stages.log 2>&1; echo "stages exit $?" | tee -a ${O}_summary.txt
grep -E "^(1|4|32|256|4096) |generate|verify" ${O}_stages.log | tee -a ${O}_summary.txt
timeout 300 python scratch/verify_timing.py > ${O}_verify.log 2>&1; echo "verify timing exit $?" | tee -a ${O}_summary.txt; tail -6 ${O}_verify.log | tee -a ${O}_summary.txt
# launch list of one 4 096-proof step (3 warm-up + 1 timed device batch), every kernel
RLN_BENCH_GLOBAL_BATCH=4096 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file ${O}_launches.csv \
    python bench.py --profile --steps 1 --warmup 3 > ${O}_ncu_list.log 2>&1; echo "ncu launch list exit $?" | tee -a ${O}_summary.txt
# full captures: the witness VM (TMA staging), both accumulate kernels, one NTT pass — 1 launch each, taken after the warm-up batches
RLN_BENCH_GLOBAL_BATCH=4096 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_witness|k_msm_accum' --launch-skip 9 -c 3 \
    -o ${O}_full python bench.py --profile --steps 1 --warmup 3 > ${O}_ncu_full.log 2>&1; echo "ncu full exit $?" | tee -a ${O}_summary.txt
ls -la ${O}_full.ncu-rep 2>/dev/null
ncu -i ${O}_full.ncu-rep --page raw --csv > ${O}_full_raw.csv 2>/dev/null; wc -c ${O}_full_raw.csv
