#!/bin/bash
# round-2 GPU job 37 (2 GPUs): the new 1-GPU tests, every multi-GPU parity test over NCCL / peer memory on the final library,
# the 2-GPU bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "swav_iteration or barlow_oracle" 2>&1 | tail -3 > gpurun_out/r2_j37_pytest_new.txt
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_j37_pytest_dist.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 2 --steps 20 --warmup 5 --no-per-config > gpurun_out/r2_j37_bench2.json 2> gpurun_out/r2_j37_bench2.err
echo done
