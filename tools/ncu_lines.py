"""Per-source-line instruction and stall-sample shares of one kernel from an ncu report (no GPU needed).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep bwd_warp [top_n]
"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k",
                          "regex:" + kern], capture_output=True, text=True).stdout
    agg, tot, tots, cur, ie, isamp = {}, 0, 0, None, None, None
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            ie, isamp = r.index("Instructions Executed"), r.index("# Samples")
            continue
        if r[0] == "Function Name" or len(r) < 8 or r[0] == "" or r[2] != "-":
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        n, sm = int(r[ie]), int(r[isamp])
        a = agg.setdefault((cur, ln, r[1][:100]), [0, 0])
        a[0] += n
        a[1] += sm
        tot += n
        tots += sm
    print("warp instructions %d, samples %d" % (tot, tots))
    top = sorted(agg.items(), key=lambda kv: -(kv[1][0] / max(tot, 1) + kv[1][1] / max(tots, 1)))[:top_n]
    for (f, ln, src), (n, sm) in sorted(top):
        print("%-16s %5d %5.1f%% ins %5.1f%% smp  %s" % (f[:16], ln, 100.0 * n / max(tot, 1), 100.0 * sm / max(tots, 1), src))


if __name__ == "__main__":
    main()
