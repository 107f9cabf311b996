#!/bin/bash
# round 2, call q: state with the lane-parallel verifier as the default up to 4 096 proofs and the witness VM's dedicated squaring:
# full GPU suite, smoke, bench line, per-stage times
set -u
mkdir -p gpurun_out
O=gpurun_out/r02q
timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "suite exit $?" | tee ${O}_summary.txt
tail -3 ${O}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke exit $?" | tee -a ${O}_summary.txt
timeout 300 python scratch/stage_breakdown.py > ${O}_stages.log 2>&1; grep -E "^(1|4|32|256|4096) |generate|verify" ${O}_stages.log | tee -a ${O}_summary.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench exit $?" | tee -a ${O}_summary.txt
tail -4 ${O}_bench.err
python - <<'PY' | tee -a gpurun_out/r02q_summary.txt
import json
l = json.loads(open('gpurun_out/r02q_bench.json').read().strip().splitlines()[-1])
print({k: l[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, l['e2e']['value'])
print(l.get('single_proof')); print(l.get('verify_batch'))
PY
