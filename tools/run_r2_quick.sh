mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2r_tests.log 2>&1; tail -3 gpurun_out/r2r_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-hr > gpurun_out/r2r_a.json 2> /dev/null
python - <<'PY'
import json
for f in "a":
    try:
        d = json.load(open(f"gpurun_out/r2r_{f}.json"))
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3),
              " ".join(f"{k}={v['ms_per_view'] and round(v['ms_per_view'],4)}" for k, v in d["kernels"].items()))
    except Exception as e:
        print(f, "FAILED", e)
PY
