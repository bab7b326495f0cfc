#!/bin/bash
# round-2 GPU job 16 (1 GPU): quick A/B of one kernel family - next-row tests + selected bench rows (ROWS env)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest ${TESTS:-tests/test_gpu_next_rows.py} -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2_j16_pytest.txt
timeout 600 python bench_losses.py --no-cpu --no-ref-gpu --reps 10 --only ${ROWS:-dino64,dino1024} --timeline gpurun_out/r2_j16_timeline.txt \
  > gpurun_out/r2_j16_losses.jsonl 2> gpurun_out/r2_j16_losses.md
echo done
