#!/bin/bash
# final-state validation: suite, smoke, bench line (N=1), launch list and full captures of the round-2 kernels
set -u
mkdir -p gpurun_out
O=gpurun_out/r02h
timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "suite exit $?" | tee ${O}_summary.txt
tail -3 ${O}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke exit $?" | tee -a ${O}_summary.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench exit $?" | tee -a ${O}_summary.txt
tail -4 ${O}_bench.err
timeout 300 python scratch/stage_breakdown.py > ${O}_stages.log 2>&1; grep -E "^(1|4|32|256|4096) |generate|verify" ${O}_stages.log | tee -a ${O}_summary.txt
RLN_BENCH_GLOBAL_BATCH=4096 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file ${O}_launches.csv \
    python bench.py --profile --steps 1 --warmup 3 > ${O}_ncu_list.log 2>&1; echo "ncu launch list exit $?" | tee -a ${O}_summary.txt
# single-proof path under ncu: tiled NTT kernels, witness VM, packed accumulate (small tables: the kernels are the same)
RLN_B200_WINDOW_BITS=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_ntt_outer|k_ntt_middle|k_witness$|k_witness\(|k_records|k_proof_records|k_witness_records' -c 8 \
    -o ${O}_full_small python scratch/single_proof.py > ${O}_ncu_small.log 2>&1; echo "ncu small exit $?" | tee -a ${O}_summary.txt
ncu -i ${O}_full_small.ncu-rep --page raw --csv > ${O}_full_small_raw.csv 2>/dev/null; wc -c ${O}_full_small_raw.csv
RLN_B200_WINDOW_BITS=8 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_single.csv python scratch/single_proof.py > /dev/null 2>&1
rm -f ${O}_full_small.ncu-rep
