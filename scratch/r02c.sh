#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c
timeout 900 python -m pytest tests -m gpu -x -q --durations=12 > ${O}_pytest.log 2>&1; echo "suite exit $?" | tee ${O}_summary.txt
tail -25 ${O}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke exit $?" | tee -a ${O}_summary.txt
timeout 300 python scratch/stage_breakdown.py > ${O}_stages.log 2>&1; echo "stages exit $?" | tee -a ${O}_summary.txt; cat ${O}_stages.log | tail -12
timeout 900 python bench.py --steps 3 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench exit $?" | tee -a ${O}_summary.txt
tail -5 ${O}_bench.err; head -c 3000 ${O}_bench.json
