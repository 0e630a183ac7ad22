"""Turns an `ncu --set full` report into a committed summary:  python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/Y.csv "comment"
(runs here, no GPU needed).  One row per captured launch: duration, DRAM bytes, DRAM / SM / tensor-pipe utilisation,
instructions, issue-active, achieved occupancy, registers, launch geometry."""
import csv, io, json, os, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
comment = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
cols = [c for c in cols if c in hdr]
idx = [hdr.index(c) for c in cols]
with open(out, "w") as f:
    for line in comment.split("\\n"):
        if line:
            f.write("# " + line + "\n")
    w = csv.writer(f)
    w.writerow(cols)
    for d in data:
        w.writerow([("%s %s" % (d[i], units[i]) if units[i] and c != "Kernel Name" else d[i]) for c, i in zip(cols, idx)])
print(open(out).read())
