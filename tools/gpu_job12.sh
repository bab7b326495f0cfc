#!/bin/bash
# round-2 GPU job 12 (1 GPU): the driver's sequence - full parity suite, smoke, default bench (both arms)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_j12_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_j12_smoke.txt 2>&1
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r2_j12_bench_ref.json 2> gpurun_out/r2_j12_bench_ref.err
( time timeout 900 python bench.py ) > gpurun_out/r2_j12_bench.json 2> gpurun_out/r2_j12_bench.err
timeout 600 python bench_losses.py --reps 20 > gpurun_out/r2_j12_losses.jsonl 2> gpurun_out/r2_j12_losses.md
echo done
