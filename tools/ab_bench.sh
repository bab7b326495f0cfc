#!/bin/bash
# A/B harness (not a test): run bench.py against every tuning build libssv_b200*.so and print kernel times.
cd "$(dirname "$0")/.."
for lib in self-supervised-vision_b200/ssv_b200/libssv_b200*.so; do
  name=$(basename $lib)
  SSVB_LIB=$name timeout 200 python bench.py --steps ${STEPS:-20} --warmup 5 --no-cpu --no-per-config 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$name', 'step %.3f ms' % d['ms_per_step'], 'bwd %.3f ms (%.1f%%)' % (r['kernel_ms'], 100*r['frac']), 'fwd %.3f ms (%.1f%%)' % (r['fwd_kernel']['ms'], 100*r['fwd_kernel']['frac']), 'step frac %.1f%%' % (100*r['step']['frac']), 'clk', d['clocks']['sm_mhz'])
"
done
