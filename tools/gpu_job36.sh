#!/bin/bash
# round-2 GPU job 36 (1 GPU): compute-sanitizer memcheck + synccheck over every kernel family incl. the third-session kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/r2_j36_memcheck.txt 2>&1
timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_small.py > gpurun_out/r2_j36_synccheck.txt 2>&1
echo done
