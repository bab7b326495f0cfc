#!/bin/bash
# round-2 GPU job 6 (2 GPUs): p2p kernel timing; on GPU 0: new parity tests (fused MoCo, ADVICE regressions), per-config bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/p2p_kernel_timing.py 2>&1 | grep -v -i "warn\|OMP\|\*\*\*" > gpurun_out/r2_j6_p2p2.txt
N=8192 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/p2p_kernel_timing.py 2>&1 | grep -v -i "warn\|OMP\|\*\*\*" > gpurun_out/r2_j6_p2p2_small.txt
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dist.py 2>&1 | tail -15 > gpurun_out/r2_j6_pytest.txt
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench_losses.py --reps 20 --no-cpu > gpurun_out/r2_j6_losses.jsonl 2> gpurun_out/r2_j6_losses.md
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/prof_host_overhead.py > gpurun_out/r2_j6_host.txt 2>&1
echo done
