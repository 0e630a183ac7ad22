mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-hr > gpurun_out/r2w_a.json 2> /dev/null
OLS_AE_L0_TF32=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-hr > gpurun_out/r2w_b.json 2> /dev/null
timeout 300 python bench.py --config 4 --steps 6 --warmup 3 --no-e2e > gpurun_out/r2w_c4.json 2> /dev/null
python - <<'PY'
import json
for f in "ab":
    d = json.load(open(f"gpurun_out/r2w_{f}.json"))
    print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "ae ms/launch", d["kernels"]["ae"]["ms_per_launch"], "frac", round(d["roofline"]["frac"],4))
d = json.load(open("gpurun_out/r2w_c4.json")); r = d["roofline"]
print("config4", d["value"], d["ms_per_step"], "enc", r["encode_ms"], r["encode_TFLOPs"], "dec", r["decode_ms"], r["decode_TFLOPs"], "two-stage", d["two_stage"]["ms_per_step"], d["parity"])
PY
