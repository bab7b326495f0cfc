#!/bin/bash
# round-2 GPU job 13 (1 GPU): diagnostics of the non-NT-Xent rows - kernel timelines of one graph replay per row and
# per-kernel DRAM metrics under ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench_losses.py --no-cpu --no-ref-gpu --reps 10 --timeline gpurun_out/r2_j13_timeline.txt \
  > gpurun_out/r2_j13_losses.jsonl 2> gpurun_out/r2_j13_losses.md
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,launch__grid_size \
  --clock-control none -c 4000 --csv --log-file gpurun_out/r2_j13_metrics.csv \
  python bench_losses.py --no-cpu --no-ref-gpu --reps 1 --only cfg2,cfg3,cfg4,swav,rowdot128,rowdot1024,relic4096,dino1024,pirl65536,ema \
  > gpurun_out/r2_j13_ncu.log 2>&1
echo done
