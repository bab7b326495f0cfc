#!/bin/bash
# round-2 GPU job 15 (1 GPU): parity suite + per-row timelines (Dino smem Tsum, wide Barlow finalize)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2_j15_pytest.txt
timeout 600 python bench_losses.py --no-cpu --no-ref-gpu --reps 10 --only cfg3,dino64,dino1024,cfg2 --timeline gpurun_out/r2_j15_timeline.txt \
  > gpurun_out/r2_j15_losses.jsonl 2> gpurun_out/r2_j15_losses.md
echo done
