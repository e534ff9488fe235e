#!/bin/bash
# usage: tools/tsweep_tier.sh <tier> workloads...
tier=$1; shift
for w in "$@"; do
  python bench.py --workload $w --tier $tier --steps 10 --warmup 3 --no-sweep --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    j=json.loads(line); r=j['roofline']
    print('%-6s %-6s step %.3f ms  %.0f seq/s | fwd %.3f ms | bwd %.3f ms' % ('$w', '$tier', j['ms_per_step'], j['value'], r['forward']['launch_ms'], r['launch_ms']))
"
done
