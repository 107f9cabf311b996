#!/bin/bash
# 8-GPU: BASELINE configs[4] exactly as the driver runs it (torchrun, global batch 65 536), then the in-process form
set -u
mkdir -p gpurun_out
O=gpurun_out/r02i
nvidia-smi -L | wc -l | tee ${O}_summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > ${O}_bench_n8.json 2> ${O}_bench_n8.err; echo "bench N=8 torchrun exit $?" | tee -a ${O}_summary.txt
tail -6 ${O}_bench_n8.err; head -c 1500 ${O}_bench_n8.json; echo
timeout 600 python bench.py --gpus 8 --inproc --steps 5 --warmup 3 > ${O}_bench_inproc8.json 2> ${O}_bench_inproc8.err; echo "bench N=8 inproc exit $?" | tee -a ${O}_summary.txt
tail -3 ${O}_bench_inproc8.err; head -c 1200 ${O}_bench_inproc8.json; echo
