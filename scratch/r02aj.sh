#!/bin/bash
# round 2, call aj (2 GPUs): the driver's N=2 launch of the final code, exactly as the contract words it (BASELINE configs[4]: global
# batch 65 536, strong scaling), the reference arm under torchrun, and the multi-device tests of the suite
set -u
mkdir -p gpurun_out
O=gpurun_out/r02aj
nvidia-smi -L | tee ${O}_summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > ${O}_bench_n2.json 2> ${O}_bench_n2.err; echo "bench N=2 torchrun exit $?" | tee -a ${O}_summary.txt
tail -4 ${O}_bench_n2.err; head -c 1200 ${O}_bench_n2.json; echo
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_device" > ${O}_pytest_multi.log 2>&1; echo "multi-device tests exit $?" | tee -a ${O}_summary.txt
tail -2 ${O}_pytest_multi.log | tee -a ${O}_summary.txt
