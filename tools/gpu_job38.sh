#!/bin/bash
# round-2 GPU job 38 (1 GPU): final sanity on the clean-built library at HEAD: the driver's GPU test command + smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_j38_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 > gpurun_out/r2_j38_smoke.txt
echo done
