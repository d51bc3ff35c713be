"""Per-kernel summary table of an .ncu-rep (raw page): duration, DRAM bytes, throughput, occupancy, issue activity.
Usage: python tools/ncu_summary.py rep.ncu-rep > profiles/xxx.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit_%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst")]
idx = [(hdr.index(m), n) for m, n in want if m in hdr]
print("| " + " | ".join(f"{n} [{units[i]}]" if units[i] else n for i, n in idx) + " |")
print("|" + "---|" * len(idx))
for r in rows[2:]:
    cells = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.split("(")[0].replace("vpc::", "")
        else:
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
