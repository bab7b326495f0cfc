#!/bin/bash
# round-2 GPU job 33 (1 GPU): SwAV with the codes rebuilt inside the cross-entropy kernel (no final Sinkhorn pass, no code matrix)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist_stages.py -m gpu -q -x -k "swav or sinkhorn or alternative" 2>&1 | tail -5 > gpurun_out/r2_j33_pytest.txt
SSVB_SWAV_NO_FUSED_CODES=1 timeout 300 python bench_losses.py --no-cpu --no-ref-gpu --reps 20 --only swav > gpurun_out/r2_j33_old.jsonl 2> gpurun_out/r2_j33_old.md
timeout 300 python bench_losses.py --no-cpu --no-ref-gpu --reps 20 --only swav,cfg4 --timeline gpurun_out/r2_j33_timeline.txt > gpurun_out/r2_j33_new.jsonl 2> gpurun_out/r2_j33_new.md
echo done
