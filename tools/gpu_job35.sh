#!/bin/bash
# round-2 GPU job 35 (1 GPU): ncu evidence of the third-session kernels: per-kernel DRAM metrics of the Barlow / Sinkhorn /
# SwAV rows, ncu --set full of the two Barlow GEMMs (forward with the dC .* C sums, dual backward with the closed-form epilogue)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "barlow" 2>&1 | tail -3 > gpurun_out/r2_j35_pytest.txt
timeout 300 python bench_losses.py --no-cpu --no-ref-gpu --reps 20 --only cfg3 > gpurun_out/r2_j35_cfg3.jsonl 2> gpurun_out/r2_j35_cfg3.md
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct \
  --clock-control none -k 'regex:^(?!vectorized|reduce|elementwise|unrolled|index|Cat|distribution).*kernel' -c 600 --csv --log-file gpurun_out/r2_j35_metrics.csv \
  python bench_losses.py --profile --no-cpu --no-ref-gpu --only cfg3,cfg4,swav > gpurun_out/r2_j35_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 2 -o gpurun_out/r2_prof_barlow_gemm_v2 -f python tools/prof_barlow.py > gpurun_out/r2_j35_ncu_full.log 2>&1
echo done
