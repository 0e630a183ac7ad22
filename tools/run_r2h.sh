mkdir -p gpurun_out
for mb in 5 4 128 1284 32; do
OLS_B2_MINB=$mb timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-hr > gpurun_out/r2h_mb$mb.json 2> /dev/null
done
python - <<'PY'
import json
for mb in (5,4,128,1284,32):
    try:
        d = json.load(open(f"gpurun_out/r2h_mb{mb}.json"))
        print(mb, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3),
              " ".join(f"{k}={v['ms_per_view'] and round(v['ms_per_view'],4)}" for k, v in d["kernels"].items()))
    except Exception as e:
        print(mb, "FAILED", e)
PY
