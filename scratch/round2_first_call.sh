#!/bin/bash
# First gpurun call of the next round: validates what round 1 committed after its GPU budget ran out, then measures the two
# experimental variants.  Usage: gpurun --timeout 1500 -- 'bash scratch/round2_first_call.sh'
# Everything lands in gpurun_out/r02a_*.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02a
# 1. the whole GPU suite as committed (includes the new V3 C caller flow and the out-of-bounds test)
timeout 900 python -m pytest tests -m gpu -q > ${O}_pytest_default.log 2>&1; echo "default suite exit $?" | tee -a ${O}_summary.txt
# 2. correctness of the experimental variants against the same oracle-backed tests
RLN_B200_WITNESS_STAGED=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "known_answer or batch_proofs or partial or multi_message or external_witness" > ${O}_pytest_staged.log 2>&1
echo "staged witness tests exit $?" | tee -a ${O}_summary.txt
RLN_B200_VARMSM_GLV=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "msm" > ${O}_pytest_varglv.log 2>&1
echo "GLV variable-base MSM tests exit $?" | tee -a ${O}_summary.txt
# 2b. request coalescing behind the single-item calls (16 threads on one handle), off and on
RLN_B200_COALESCE=0 timeout 300 python scratch/coalesce_gpu_check.py > ${O}_coalesce_off.log 2>&1; echo "coalesce off exit $? $(tail -1 ${O}_coalesce_off.log)" | tee -a ${O}_summary.txt
RLN_B200_COALESCE=1 timeout 300 python scratch/coalesce_gpu_check.py > ${O}_coalesce_on.log 2>&1; echo "coalesce on exit $? $(tail -1 ${O}_coalesce_on.log)" | tee -a ${O}_summary.txt
RLN_B200_COALESCE=1 timeout 900 python -m pytest tests -m gpu -q > ${O}_pytest_coalesce.log 2>&1; echo "suite with coalescing exit $?" | tee -a ${O}_summary.txt
# 3. numbers: default, staged witness, GLV variable-base MSM (the sweep and single-proof latency are in the bench line)
timeout 400 python bench.py --steps 3 --warmup 3 > ${O}_bench_default.json 2> ${O}_bench_default.err
RLN_B200_WITNESS_STAGED=1 timeout 400 python bench.py --steps 3 --warmup 3 > ${O}_bench_staged.json 2> ${O}_bench_staged.err
RLN_B200_VARMSM_GLV=1 timeout 400 python bench.py --steps 3 --warmup 3 > ${O}_bench_varglv.json 2> ${O}_bench_varglv.err
python - <<'PY' | tee -a gpurun_out/r02a_summary.txt
import json
for tag in ("default", "staged", "varglv"):
    try:
        d = json.loads(open(f"gpurun_out/r02a_bench_{tag}.json").read().strip().splitlines()[-1])
        print(tag, round(d["value"], 1), "proofs/s", "witness", round(d["stage_ms"]["witness"], 2), "ms", "single", d.get("single_proof"),
              "msm sweep", [(r["log2_n"], round(r["ms"], 2)) for r in d.get("msm_g1_sweep", [])])
    except Exception as e:
        print(tag, "no bench line:", e)
PY
tail -3 ${O}_pytest_default.log ${O}_pytest_staged.log ${O}_pytest_varglv.log
