"""Print the headline fields of a bench.py JSON line (file argument or stdin)."""
import json, sys
txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
d = json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
r = d.get('roofline', {})
print('%s | value %.0f %s | %.4f ms/step | e2e %.0f | bwd %.4f ms frac %.4f | fwd %.4f ms frac %.4f | launches %s | clocks %s' % (
    d['config'].get('workload', '?')[:30], d['value'], d['unit'], d['ms_per_step'], d.get('e2e', {}).get('value', 0),
    r.get('launch_ms', 0), r.get('frac', 0), r.get('forward', {}).get('launch_ms', 0), r.get('forward', {}).get('frac', 0),
    d.get('gpu_launches'), d.get('clocks', {}).get('sm_mhz')))
for v in (r.get('sweep') or []):
    print('   sweep', v)
