#!/bin/bash
# round-2 GPU job 27 (1 GPU): per-kernel DRAM metrics (ncu) of the final kernels on selected rows, incl. the beyond-L2 shapes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct \
  --clock-control none -k 'regex:^(?!vectorized|reduce|elementwise|unrolled|index|Cat|distribution).*kernel' -c 1500 --csv --log-file gpurun_out/r2_j27_metrics.csv \
  python bench_losses.py --profile --no-cpu --no-ref-gpu --only cfg2,cfg3,cfg4,swav,rowdot128,rowdot_big,dino1024,pirl65536,ema,ema_big \
  > gpurun_out/r2_j27_ncu.log 2>&1
echo done
