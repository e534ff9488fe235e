"""Text summary of an ncu report for profiles/: per kernel launch the duration, instruction counts, IPC, occupancy,
top stall reasons, FMA-pipe utilisation and DRAM bytes (no GPU needed).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>_ncu_summary.txt
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC per SM (active)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % of peak (inst)"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active %"),
    ("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "FFMA thread-instructions"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    print("# %s" % rep.split("/")[-1])
    for r in rows[2:]:
        print("\n== %s" % r[name_i])
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-32s %s %s" % (label, r[i], units[i]))
        st = sorted(((float(r[i] or 0), h) for i, h in stall), reverse=True)[:6]
        print("  top stalls (warps per issue): " + ", ".join("%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, h in st))


if __name__ == "__main__":
    main()
