#!/bin/bash
# round-2 GPU job 34 (1 GPU): the driver's sequence on the final library - full GPU test suite, smoke, default bench
# (both arms), full per-loss table with timelines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_j34_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_j34_smoke.txt 2>&1
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r2_j34_bench_ref.json 2> gpurun_out/r2_j34_bench_ref.err
( time timeout 900 python bench.py ) > gpurun_out/r2_j34_bench.json 2> gpurun_out/r2_j34_bench.err
timeout 900 python bench_losses.py --reps 20 --timeline gpurun_out/r2_j34_timeline.txt > gpurun_out/r2_j34_losses.jsonl 2> gpurun_out/r2_j34_losses.md
echo done
