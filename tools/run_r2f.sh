set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2f_tests.log 2>&1; tail -12 gpurun_out/r2f_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-hr > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -3 gpurun_out/r2f_bench.err
OLS_BWD_V1=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-hr > gpurun_out/r2f_bench_bwdv1.json 2> /dev/null
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-hr --tile 16 --backward-mode exact > gpurun_out/r2f_bench_exact16.json 2> /dev/null
OLS_BWD_V1=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-hr --tile 16 --backward-mode exact > gpurun_out/r2f_bench_exact16_v1.json 2> /dev/null
timeout 600 python bench.py --config 5 --steps 8 --warmup 1 > gpurun_out/r2f_config5.json 2> gpurun_out/r2f_config5.err; tail -3 gpurun_out/r2f_config5.err
python - <<'PY'
import json
for f in ("r2f_bench", "r2f_bench_bwdv1", "r2f_bench_exact16", "r2f_bench_exact16_v1"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3),
              " ".join(f"{k}={v['ms_per_view'] and round(v['ms_per_view'],4)}" for k, v in d["kernels"].items()))
    except Exception as e:
        print(f, "FAILED", e)
try:
    d = json.load(open("gpurun_out/r2f_config5.json"))
    print("config5 fps", d["value"], "track ms/it", d["tracking_ms_per_iteration"], "map ms/it", d["mapping_ms_per_iteration"], "ae ms/kf", d["ae_ms_per_keyframe"])
except Exception as e:
    print("config5 FAILED", e)
PY
