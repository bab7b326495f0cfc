#!/bin/bash
# round-2 GPU job 4 (2 GPUs): multi-GPU parity over NCCL + the new peer-memory flag transport, 2-GPU bench, stage timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_j4_topo.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_dist.py tests/test_gpu_dist_stages.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2_j4_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_j4_bench2.json 2> gpurun_out/r2_j4_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_stage_timing.py > gpurun_out/r2_j4_stage2.txt 2>&1
echo done
