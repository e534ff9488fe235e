#!/bin/bash
# T-sweep of the block tier (kernel times + FP32 fractions) -> gpurun_out/tsweep.txt
for w in "$@"; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-sweep --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    j=json.loads(line); r=j['roofline']
    print('%-6s step %.3f ms  %.0f seq/s | fwd %.3f ms frac %.3f (model %.3f) | bwd %.3f ms frac %.3f (model %.3f)' % ('$w', j['ms_per_step'], j['value'], r['forward']['launch_ms'], r['forward']['frac'], r['forward']['model_frac'], r['launch_ms'], r['frac'], r['model_frac']))
"
done
