#!/bin/bash
# round-2 GPU job 8 (1 GPU): full parity suite, default bench, launch list + ncu --set full of both similarity kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_j8_pytest.txt
STEPS=20 timeout 300 bash tools/ab_bench.sh > gpurun_out/r2_j8_ab.txt 2>&1
STEPS=20 timeout 300 bash tools/ab_bench.sh >> gpurun_out/r2_j8_ab.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sim_bwd_kernel -s 2 -c 1 -o gpurun_out/r2_prof_bwd_j8 -f python tools/prof_ntxent.py > gpurun_out/r2_j8_ncu_bwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sim_fwd_kernel -s 2 -c 1 -o gpurun_out/r2_prof_fwd_j8 -f python tools/prof_ntxent.py > gpurun_out/r2_j8_ncu_fwd.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_j8_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-per-config > gpurun_out/r2_j8_bench_under_ncu.log 2>&1
timeout 300 python tools/prof_host_overhead.py 2>&1 | head -12 > gpurun_out/r2_j8_host.txt
echo done
