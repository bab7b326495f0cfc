#!/bin/bash
# round-2 GPU job 18 (1 GPU): full parity suite (incl. the new 128 < d <= 256 cases), host-overhead profile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_j18_pytest.txt
timeout 300 python tools/prof_host_overhead.py > gpurun_out/r2_j18_host.txt 2>&1
echo done
