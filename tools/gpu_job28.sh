#!/bin/bash
# round-2 GPU job 28 (1 GPU): full parity suite on the final library + host-us of a few rows (tensor-map cache)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_j28_pytest.txt
timeout 300 python bench_losses.py --no-cpu --no-ref-gpu --reps 10 --only cfg1,cfg2,cfg3,relic512 > gpurun_out/r2_j28.jsonl 2> gpurun_out/r2_j28.md
echo done
