#!/bin/bash
# round-2 GPU job 26 (1 GPU): Sinkhorn passes behind the alpha kernels with programmatic dependent launch: parity + A/B timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next_rows.py tests/test_gpu_dist_stages.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2_j26_pytest.txt
SSVB_NO_PDL=1 timeout 300 python bench_losses.py --no-cpu --no-ref-gpu --reps 20 --only cfg4,swav > gpurun_out/r2_j26_nopdl.jsonl 2> gpurun_out/r2_j26_nopdl.md
timeout 300 python bench_losses.py --no-cpu --no-ref-gpu --reps 20 --only cfg4,swav --timeline gpurun_out/r2_j26_timeline.txt > gpurun_out/r2_j26_pdl.jsonl 2> gpurun_out/r2_j26_pdl.md
echo done
