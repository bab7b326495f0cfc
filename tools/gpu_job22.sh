#!/bin/bash
# round-2 GPU job 22 (2 GPUs): every multi-GPU parity test over NCCL / peer memory, then the 2-GPU bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_j22_pytest_dist.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_j22_bench2.json 2> gpurun_out/r2_j22_bench2.err
echo done
