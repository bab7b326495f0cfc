#!/bin/bash
# round-2 GPU job 17 (1 GPU): A/B of the L2 flush method (write-only vs write + read-back) on selected rows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ROWS=${ROWS:-cfg2,cfg3,cfg4,rowdot128,dino1024,swav}
#BENCH_FLUSH=write timeout 600 python bench_losses.py --no-cpu --no-ref-gpu --reps 10 --only $ROWS --timeline gpurun_out/r2_j17_timeline_write.txt > gpurun_out/r2_j17_write.jsonl 2> gpurun_out/r2_j17_write.md
BENCH_FLUSH=clean timeout 600 python bench_losses.py --no-cpu --no-ref-gpu --reps 10 --only $ROWS --timeline gpurun_out/r2_j17_timeline_clean.txt > gpurun_out/r2_j17_clean.jsonl 2> gpurun_out/r2_j17_clean.md
echo done
