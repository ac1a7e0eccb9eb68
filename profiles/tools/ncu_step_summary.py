"""Per-kernel summary of an `ncu --set full` capture of one per-sample step (the CSV kept under profiles/): duration, DRAM bytes,
warp instructions, issue-slot utilisation, occupancy, launch shape.  usage: ncu_step_summary.py in.ncu-rep out.csv"""
import csv
import io
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__grid_size", "launch__block_size", "launch__registers_per_thread"]
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
head, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(head)}
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["#", "kernel"] + ["%s [%s]" % (m, units[col[m]]) if units[col[m]] else m for m in METRICS])
    tot_us = tot_mb = 0.0
    for i, r in enumerate(data):
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("unnamed>::", "")
        w.writerow([i, name] + [r[col[m]] for m in METRICS])
        tot_us += float(r[col["gpu__time_duration.sum"]])
        tot_mb += float(r[col["dram__bytes_read.sum"]]) + float(r[col["dram__bytes_write.sum"]])
    w.writerow(["sum", "%d kernels" % len(data), "%.3f" % tot_us, "%.3f (read + write, MB)" % tot_mb])
print("%d kernels, %.1f us, %.1f MB of DRAM traffic" % (len(data), tot_us, tot_mb))
