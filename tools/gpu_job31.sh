#!/bin/bash
# round-2 GPU job 31 (1 GPU): Barlow closed-form backward in the GEMM epilogue + SwAV backward (one memset / one scale launch)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist_stages.py -m gpu -q -x -k "swav or barlow or alternative" 2>&1 | tail -5 > gpurun_out/r2_j31_pytest.txt
timeout 120 python tools/barlow_err.py > gpurun_out/r2_j31_err_fused.txt 2>&1
SSVB_BARLOW_NO_FUSED_BWD=1 timeout 120 python tools/barlow_err.py > gpurun_out/r2_j31_err_unfused.txt 2>&1
SSVB_BARLOW_NO_FUSED_BWD=1 timeout 300 python bench_losses.py --no-cpu --no-ref-gpu --reps 20 --only cfg3 > gpurun_out/r2_j31_old.jsonl 2> gpurun_out/r2_j31_old.md
timeout 300 python bench_losses.py --no-cpu --no-ref-gpu --reps 20 --only cfg3,swav --timeline gpurun_out/r2_j31_timeline.txt > gpurun_out/r2_j31_new.jsonl 2> gpurun_out/r2_j31_new.md
echo done
