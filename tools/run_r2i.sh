set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2i_tests.log 2>&1; tail -12 gpurun_out/r2i_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-hr > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -3 gpurun_out/r2i_bench.err
python - <<'PY'
import json
for f in ("r2i_bench",):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3),
              " ".join(f"{k}={v['ms_per_view'] and round(v['ms_per_view'],4)}" for k, v in d["kernels"].items()))
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_' -s 66 -c 22 --csv --log-file gpurun_out/r2i_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hr --no-graph > /dev/null 2> gpurun_out/r2i_ncu.err
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2i_launches.csv")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
kn, mn, mv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    key = (r[h.index("ID")], r[kn][:40])
    agg.setdefault(key, {})[r[mn]] = r[mv]
for (i, n), m in agg.items():
    print(n.ljust(42), m)
PY
