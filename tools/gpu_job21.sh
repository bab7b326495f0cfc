#!/bin/bash
# round-2 GPU job 21 (1 GPU): the driver's sequence - smoke, default bench (both arms), full per-loss table with timelines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_j21_smoke.txt 2>&1
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r2_j21_bench_ref.json 2> gpurun_out/r2_j21_bench_ref.err
( time timeout 900 python bench.py ) > gpurun_out/r2_j21_bench.json 2> gpurun_out/r2_j21_bench.err
timeout 900 python bench_losses.py --reps 20 --timeline gpurun_out/r2_j21_timeline.txt > gpurun_out/r2_j21_losses.jsonl 2> gpurun_out/r2_j21_losses.md
echo done
