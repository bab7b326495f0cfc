#!/bin/bash
# round-2 GPU job 5 (8 GPUs): NT-Xent multi-GPU parity (world 4 inside the tests), stage timing and bench at 8 / 4 GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x -k "ntxent or relic" 2>&1 | tail -8 > gpurun_out/r2_j5_pytest.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/dist_stage_timing.py 2>&1 | grep -v -i "warn\|OMP\|\*\*\*" > gpurun_out/r2_j5_stage8.txt
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_j5_bench$n.json 2> gpurun_out/r2_j5_bench$n.err
done
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-per-config > gpurun_out/r2_j5_bench1.json 2> gpurun_out/r2_j5_bench1.err
echo done
