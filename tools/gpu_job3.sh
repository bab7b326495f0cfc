#!/bin/bash
# round-2 GPU job 3 (1 GPU): new parity tests (emulated-rank NT-Xent stages + p2p flags transport, gradient witness at
# N = 32768), full default bench.py line (parity + per_config + reference CPU arm)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2_j3_pytest.txt
( time timeout 900 python bench.py ) > gpurun_out/r2_j3_bench.json 2> gpurun_out/r2_j3_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2_j3_bench_ref.json 2> gpurun_out/r2_j3_bench_ref.err
echo done
