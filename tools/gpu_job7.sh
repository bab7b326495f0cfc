#!/bin/bash
# round-2 GPU job 7 (1 GPU): parity with the 4-buffer forward kernel + fused losses, A/B vs the 256-wide forward, host profile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dist.py 2>&1 | tail -15 > gpurun_out/r2_j7_pytest.txt
STEPS=20 timeout 600 bash tools/ab_bench.sh > gpurun_out/r2_j7_ab.txt 2>&1
STEPS=20 timeout 600 bash tools/ab_bench.sh >> gpurun_out/r2_j7_ab.txt 2>&1
timeout 300 python tools/prof_host_overhead.py > gpurun_out/r2_j7_host.txt 2>&1
timeout 300 python bench_losses.py --reps 20 --no-cpu --no-ref-gpu --only cfg1,ntx2048,ntx8192,cfg2,relic4096 > gpurun_out/r2_j7_losses.jsonl 2> gpurun_out/r2_j7_losses.md
echo done
