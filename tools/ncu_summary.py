"""Turns gpurun_out/r01_full.ncu-rep into profiles/r01_ncu_full_summary.csv and profiles/blend_fwd_dram_bytes.json.
Run here (no GPU needed):  python tools/ncu_summary.py"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(ROOT, "gpurun_out", "r01_full.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "Grid Size", "Block Size"]
idx = [hdr.index(c) for c in cols]
out = os.path.join(ROOT, "profiles", "r01_ncu_full_summary.csv")
with open(out, "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:'k_blend|k_ae_chain' -s 30 -c 3  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hr --no-graph\n")
    f.write("# (tools/run_artifacts.sh; P=1M, F=15, 960x540, 15x15 tiles, compat backward).  raw report: gpurun_out/r01_full.ncu-rep (not committed)\n")
    f.write("# per-launch values of one keyframe: AE encode, forward blend, backward blend (packed compat variant, 128 threads per tile)\n")
    w = csv.writer(f)
    w.writerow(cols)
    for d in data:
        w.writerow(["%s %s" % (d[i], units[i]) if units[i] and c != "Kernel Name" else d[i] for c, i in zip(cols, idx)])
for d in data:
    if d[idx[0]].startswith("k_blend<") or "k_blend<" in d[idx[0]] and "bwd" not in d[idx[0]]:
        def to_bytes(v, u):
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        tot = to_bytes(d[idx[2]], units[idx[2]]) + to_bytes(d[idx[3]], units[idx[3]])
        json.dump({"kernel": "k_blend<15,3,15,0>", "dram_bytes_per_launch": tot,
                   "source": "profiles/r01_ncu_full_summary.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)",
                   "note": "inputs are L2-warm (records written by the preceding kernels fit the 126 MB L2); the kernel traverses only the first ~170-300 of ~2400 entries of a tile's list"},
                  open(os.path.join(ROOT, "profiles", "blend_fwd_dram_bytes.json"), "w"), indent=1)
        print("blend fwd dram bytes", tot)
        break
print(open(out).read())
