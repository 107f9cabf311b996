#!/bin/bash
# round 2, call as: validation of the final code: full GPU suite (with the reference's six C programs), smoke, bench line (N = 1)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02as
timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "suite exit $?" | tee ${O}_summary.txt
tail -3 ${O}_pytest.log | tee -a ${O}_summary.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke exit $?" | tee -a ${O}_summary.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench exit $?" | tee -a ${O}_summary.txt
tail -3 ${O}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > ${O}_bench_reference.json 2> ${O}_bench_reference.err; echo "reference arm exit $?" | tee -a ${O}_summary.txt
head -c 600 ${O}_bench_reference.json
