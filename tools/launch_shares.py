"""Kernel shares of GPU time from an `ncu --metrics gpu__time_duration.sum --csv` launch list (no GPU needed).

    python tools/launch_shares.py gpurun_out/c2_launches.csv [skip_first_n] > profiles/rNN_c2_launch_shares.txt
"""
import csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = {}
for r in rows[skip:]:
    if "fp32_peak" in r[ik]:
        continue
    a = agg.setdefault(r[ik], [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", "")) / 1e3
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-90s n=%3d total=%10.1f us share=%5.1f%% avg=%8.1f us" % (k[-90:], n, t, 100 * t / tot, t / n))
