#!/bin/bash
# round-2 GPU job 19 (1 GPU): phase breakdown of the fused MoCo kernel (debug-timing build)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SSVB_LIB=libssv_b200_dbg.so timeout 300 python tools/dbg_moco.py > gpurun_out/r2_j19_moco_dbg.txt 2>&1
echo done
