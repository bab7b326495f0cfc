#!/bin/bash
# round-2 GPU job 39 (1 GPU): after the epilogue-loop pragma fix / dead-code removal: Barlow + SwAV parity and timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist_stages.py -m gpu -q -x -k "barlow or swav or alternative" 2>&1 | tail -3 > gpurun_out/r2_j39_pytest.txt
timeout 300 python bench_losses.py --no-cpu --no-ref-gpu --reps 20 --only cfg3,swav --timeline gpurun_out/r2_j39_timeline.txt > gpurun_out/r2_j39.jsonl 2> gpurun_out/r2_j39.md
echo done
