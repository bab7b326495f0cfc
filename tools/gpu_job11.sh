#!/bin/bash
# round-2 GPU job 11 (8 GPUs): scaling record on ONE box: bench at 8 / 4 / 2 / 1 GPUs + rank-0 timeline of the 8-GPU step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 8 4 2; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_j11_bench$n.json 2> gpurun_out/r2_j11_bench$n.err
done
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-per-config > gpurun_out/r2_j11_bench1.json 2> gpurun_out/r2_j11_bench1.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/dist_timeline.py 2>&1 | grep -v -i "warn\|OMP\|\*\*\*" > gpurun_out/r2_j11_timeline8.txt
echo done
