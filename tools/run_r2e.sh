set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2e_launches.csv \
   python bench.py --config 5 --steps 1 --warmup 0 --tracking-iters 3 --mapping-iters 2 --no-graph > /dev/null 2> gpurun_out/r2e.err; tail -3 gpurun_out/r2e.err
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2e_launches.csv")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
seq = [(r[kn], float(r[mv].replace(",", ""))) for r in rows[hi + 2:] if len(r) > mv]
# last ~400 launches = the timed frame (tracking 3 its, then keyframe mapping 2 its)
tail = seq[-420:]
agg = collections.defaultdict(lambda: [0, 0.0])
for n, t in tail:
    agg[n[:70]][0] += 1; agg[n[:70]][1] += t
tot = sum(v[1] for v in agg.values())
print("total us", tot / 1e3)
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{t/1e3:10.1f} us  x{c:4d}  {t/c/1e3:8.1f} us/launch  {n}")
PY
