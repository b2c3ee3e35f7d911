"""Condense `ncu -i rep --page raw --csv` into one line per launch: duration, DRAM bytes, issue / occupancy, registers,
executed warp instructions.  usage: python tools/ncu_summary.py raw.csv > summary.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
cols = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
idx = [hdr.index(c) for c in cols if c in hdr]
w = csv.writer(sys.stdout)
w.writerow([f"{hdr[i]} [{units[i]}]" if units[i] else hdr[i] for i in idx])
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    out = [r[i] for i in idx]
    out[0] = out[0].replace("void (anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("void <unnamed>::", "").replace("<unnamed>::", "")
    w.writerow(out)
