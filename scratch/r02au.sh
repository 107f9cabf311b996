#!/bin/bash
# round 2, call au: proof-values side kernel forked after the witness kernel (the two could not share SMs); validation of the final code
set -u
mkdir -p gpurun_out
O=gpurun_out/r02au
timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "suite exit $?" | tee ${O}_summary.txt
tail -3 ${O}_pytest.log | tee -a ${O}_summary.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke exit $?" | tee -a ${O}_summary.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench exit $?" | tee -a ${O}_summary.txt
tail -2 ${O}_bench.err
python -c "
import json;d=json.loads(open('${O}_bench.json').read().strip().splitlines()[-1]);print(d['value'], d['e2e']['value'], d['stage_ms_per_device_batch'])" | tee -a ${O}_summary.txt
