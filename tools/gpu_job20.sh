#!/bin/bash
# round-2 GPU job 20 (1 GPU): ncu --set full of the Barlow GEMMs (fused-loss forward, dual backward with TMA-store epilogue)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 2 -o gpurun_out/r2_prof_barlow_gemm -f python tools/prof_barlow.py > gpurun_out/r2_j20_ncu.log 2>&1
echo done
