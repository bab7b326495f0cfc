#!/bin/bash
# round-2 GPU job 9 (2 GPUs): per-kernel timeline of the distributed NT-Xent step (rank 0), full-size and 8-GPU-like per-rank size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_timeline.py 2>&1 | grep -v -i "warn\|OMP\|\*\*\*" > gpurun_out/r2_j9_timeline2.txt
N=8192 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dist_timeline.py 2>&1 | grep -v -i "warn\|OMP\|\*\*\*" > gpurun_out/r2_j9_timeline2_small.txt
echo done
