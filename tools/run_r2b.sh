set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2b_tests.log 2>&1; tail -15 gpurun_out/r2b_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-hr > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -3 gpurun_out/r2b_bench.err
timeout 300 python bench.py --config 4 --steps 6 --warmup 3 --no-e2e > gpurun_out/r2b_config4.json 2> gpurun_out/r2b_config4.err; tail -3 gpurun_out/r2b_config4.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2b_bench.json"))
    print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3), "launches/step", d.get("gpu_launches_per_step"))
    for k, v in d["kernels"].items():
        print("   ", k, round(v["ms_per_launch"], 4), v.get("ms_per_view") and round(v["ms_per_view"], 4))
except Exception as e:
    print("bench FAILED", e)
try:
    d = json.load(open("gpurun_out/r2b_config4.json"))
    r = d["roofline"]
    print("config4 maps/s", d["value"], "ms", d["ms_per_step"], "enc", r["encode_ms"], r["encode_TFLOPs"], "dec", r["decode_ms"], r["decode_TFLOPs"], r["decode_GBps"], d["two_stage"], d["parity"])
except Exception as e:
    print("config4 FAILED", e)
PY
