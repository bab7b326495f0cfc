#!/bin/bash
# round-2 GPU job 30 (1 GPU): two-view launches (Sinkhorn batching in SwAV, Barlow pre-pass / finish), fused SwAV staging,
# float4 SwAV cross-entropy: parity + A/B timing by the env switches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist_stages.py -m gpu -q -x -k "swav or sinkhorn or barlow or alternative" 2>&1 | tail -5 > gpurun_out/r2_j30_pytest.txt
SSVB_SK_NO_BATCH=1 SSVB_SWAV_NO_CE4=1 SSVB_BARLOW_NO_X2=1 timeout 300 python bench_losses.py --no-cpu --no-ref-gpu --reps 20 --only cfg3,cfg4,swav > gpurun_out/r2_j30_old.jsonl 2> gpurun_out/r2_j30_old.md
timeout 300 python bench_losses.py --no-cpu --no-ref-gpu --reps 20 --only cfg3,cfg4,swav --timeline gpurun_out/r2_j30_timeline.txt > gpurun_out/r2_j30_new.jsonl 2> gpurun_out/r2_j30_new.md
echo done
