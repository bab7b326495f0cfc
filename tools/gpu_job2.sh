#!/bin/bash
# round-2 GPU job 2: A/B of the polynomial shares (MOD/CNT knobs), first run of the per-config bench with reference arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STEPS=20 timeout 900 bash tools/ab_bench.sh > gpurun_out/r2_j2_ab.txt 2>&1
timeout 600 python bench_losses.py --reps 10 > gpurun_out/r2_j2_losses.jsonl 2> gpurun_out/r2_j2_losses.md
STEPS=20 timeout 900 bash tools/ab_bench.sh >> gpurun_out/r2_j2_ab.txt 2>&1
echo done
