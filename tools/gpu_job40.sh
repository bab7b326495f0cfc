#!/bin/bash
# round-2 GPU job 40 (1 GPU): wider Barlow finalize kernels: parity + timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "barlow" 2>&1 | tail -2 > gpurun_out/r2_j40_pytest.txt
timeout 40 python bench_losses.py --no-cpu --no-ref-gpu --reps 20 --only cfg3 --timeline gpurun_out/r2_j40_timeline.txt > gpurun_out/r2_j40.jsonl 2> gpurun_out/r2_j40.md
echo done
