#!/bin/bash
# 2-GPU call: full suite (multi-device tests see two devices), witness mapping A/B, torchrun bench at N=2 (small global batch), in-process bench
set -u
mkdir -p gpurun_out
O=gpurun_out/r02d
nvidia-smi -L | tee ${O}_summary.txt
timeout 1200 python -m pytest tests -m gpu -x -q --durations=12 > ${O}_pytest.log 2>&1; echo "suite exit $?" | tee -a ${O}_summary.txt
tail -30 ${O}_pytest.log
for m in 0 1; do
  RLN_B200_WITNESS_WARP=$m timeout 300 python scratch/stage_breakdown.py > ${O}_stages_warp$m.log 2>&1; echo "stages warp=$m exit $?" | tee -a ${O}_summary.txt
  grep -E "^(1|4|32|256|4096) |generate|verify" ${O}_stages_warp$m.log | tee -a ${O}_summary.txt
done
RLN_BENCH_GLOBAL_BATCH=16384 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > ${O}_bench_n2.json 2> ${O}_bench_n2.err; echo "bench N=2 torchrun exit $?" | tee -a ${O}_summary.txt
tail -4 ${O}_bench_n2.err; head -c 1800 ${O}_bench_n2.json; echo
RLN_BENCH_GLOBAL_BATCH=16384 timeout 900 python bench.py --gpus 2 --inproc --steps 3 --warmup 3 > ${O}_bench_inproc2.json 2> ${O}_bench_inproc2.err; echo "bench N=2 inproc exit $?" | tee -a ${O}_summary.txt
tail -3 ${O}_bench_inproc2.err; head -c 1500 ${O}_bench_inproc2.json; echo
