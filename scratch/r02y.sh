#!/bin/bash
# round 2, call y: validation of the final code: GPU suite, smoke, bench line (N = 1), per-stage times, verifier timings, ncu launch
# lists (device batch of 4 096; single proof + single verification), compute-sanitizer memcheck / racecheck over every kernel family
set -u
mkdir -p gpurun_out
O=gpurun_out/r02y
timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "suite exit $?" | tee ${O}_summary.txt
tail -3 ${O}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke exit $?" | tee -a ${O}_summary.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench exit $?" | tee -a ${O}_summary.txt
tail -3 ${O}_bench.err
timeout 300 python scratch/stage_breakdown.py > ${O}_stages.log 2>&1; grep -E "^(1|4|32|256|4096) |generate|verify" ${O}_stages.log | tee -a ${O}_summary.txt
timeout 600 python scratch/verify_timing.py > ${O}_verify.log 2>&1; grep -E "ms|program" ${O}_verify.log | tee -a ${O}_summary.txt
timeout 300 python scratch/vm_trace.py > ${O}_vm_trace.txt 2>&1; echo "vm trace exit $?" | tee -a ${O}_summary.txt
RLN_BENCH_GLOBAL_BATCH=4096 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file ${O}_launches.csv \
    python bench.py --profile --steps 1 --warmup 3 > ${O}_ncu_list.log 2>&1; echo "ncu launch list exit $?" | tee -a ${O}_summary.txt
RLN_B200_WINDOW_BITS=8 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_single.csv python scratch/single_proof.py > /dev/null 2>&1; echo "ncu single exit $?" | tee -a ${O}_summary.txt
timeout 900 compute-sanitizer --tool memcheck python scratch/sanitize.py > ${O}_memcheck.log 2>&1; echo "memcheck exit $?" | tee -a ${O}_summary.txt
tail -3 ${O}_memcheck.log | tee -a ${O}_summary.txt
timeout 1500 compute-sanitizer --tool racecheck python scratch/sanitize.py > ${O}_racecheck.log 2>&1; echo "racecheck exit $?" | tee -a ${O}_summary.txt
tail -3 ${O}_racecheck.log | tee -a ${O}_summary.txt
