#!/bin/bash
# round-2 GPU job 10 (2 GPUs): multi-GPU parity after the single-call forward, timeline, 2-GPU bench; previously failing 1-GPU tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next_rows.py -m gpu -q 2>&1 | tail -12 > gpurun_out/r2_j10_pytest1.txt
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x -k "ntxent or relic" 2>&1 | tail -8 > gpurun_out/r2_j10_pytest2.txt
N=8192 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dist_timeline.py 2>&1 | grep -v -i "warn\|OMP\|\*\*\*" > gpurun_out/r2_j10_timeline2_small.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_j10_bench2.json 2> gpurun_out/r2_j10_bench2.err
echo done
