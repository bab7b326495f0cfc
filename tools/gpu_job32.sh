#!/bin/bash
# round-2 GPU job 32 (1 GPU): Barlow closed-form backward with dC-consistent sums: error vs the unfused path, timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python tools/barlow_err.py > gpurun_out/r2_j32_err_fused.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist_stages.py -m gpu -q -x -k "barlow or alternative" 2>&1 | tail -5 > gpurun_out/r2_j32_pytest.txt
timeout 300 python bench_losses.py --no-cpu --no-ref-gpu --reps 20 --only cfg3 > gpurun_out/r2_j32_new.jsonl 2> gpurun_out/r2_j32_new.md
echo done
