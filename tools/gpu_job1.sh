#!/bin/bash
# round-2 GPU job 1: parity after the prescale / packed-poly / LDS changes, A/B of the polynomial shares, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_j1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dist.py 2>&1 | tail -25 > gpurun_out/r2_j1_pytest.txt
STEPS=20 timeout 600 bash tools/ab_bench.sh > gpurun_out/r2_j1_ab.txt 2>&1
STEPS=20 timeout 600 bash tools/ab_bench.sh >> gpurun_out/r2_j1_ab.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sim_bwd_kernel -s 2 -c 1 -o gpurun_out/r2_prof_bwd_j1 -f python tools/prof_ntxent.py > gpurun_out/r2_j1_ncu_bwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sim_fwd_kernel -s 2 -c 1 -o gpurun_out/r2_prof_fwd_j1 -f python tools/prof_ntxent.py > gpurun_out/r2_j1_ncu_fwd.log 2>&1
echo done
