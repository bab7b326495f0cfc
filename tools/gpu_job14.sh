#!/bin/bash
# round-2 GPU job 14 (1 GPU): parity suite on the new GEMM epilogue / fused kernels, then per-row timelines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2_j14_pytest.txt
SSVB_GEMM_NO_TMA_STORE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "barlow or swav" 2>&1 | tail -8 > gpurun_out/r2_j14_pytest_fallback.txt
timeout 600 python bench_losses.py --no-cpu --no-ref-gpu --reps 10 --timeline gpurun_out/r2_j14_timeline.txt \
  > gpurun_out/r2_j14_losses.jsonl 2> gpurun_out/r2_j14_losses.md
echo done
