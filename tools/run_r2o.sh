set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err; tail -4 gpurun_out/r2o_bench_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2o_bench_n1.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "launches/step", d["gpu_launches_per_step"])
e = d["e2e"]
print("e2e", e["value"], e["ms_per_step"], e["h2d_GBps_achieved"], e.get("h2d_ceiling"), e.get("transfer_floor_ms_per_step"))
h = e.get("hr_input_variant")
print("e2e hr variant", h and (h["value"], h["ms_per_step"], h["h2d_bytes_per_step"]))
print("cpu", d["cpu_baseline"]); print("hr", d["hr_module"] and d["hr_module"]["ms_per_frame"]); print("clocks", d["clocks"])
PY
