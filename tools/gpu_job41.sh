#!/bin/bash
# round-2 GPU job 41 (1 GPU): the driver's GPU test command on the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 45 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2 > gpurun_out/r2_j41_pytest.txt
echo done
