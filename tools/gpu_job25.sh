#!/bin/bash
# round-2 GPU job 25 (1 GPU): compute-sanitizer memcheck + synccheck over every kernel family on small shapes; the new tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/r2_j25_memcheck.txt 2>&1
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_small.py > gpurun_out/r2_j25_synccheck.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "swav or gemm" 2>&1 | tail -3 > gpurun_out/r2_j25_pytest.txt
echo done
