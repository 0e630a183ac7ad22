# round-2 artefacts on one GPU (copied into profiles/ afterwards); every step under its own timeout
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_tests.log 2>&1; tail -3 gpurun_out/r02_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -3 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -3 gpurun_out/r02_bench_reference.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-hr --per-view > gpurun_out/r02_bench_perview.json 2> /dev/null
timeout 300 python bench.py --config 4 --steps 6 --warmup 3 > gpurun_out/r02_config4.json 2> gpurun_out/r02_config4.err; tail -2 gpurun_out/r02_config4.err
timeout 600 python bench.py --config 5 --steps 8 --warmup 1 > gpurun_out/r02_config5_n1.json 2> gpurun_out/r02_config5_n1.err; tail -2 gpurun_out/r02_config5_n1.err
timeout 600 python bench.py --config 5 --steps 8 --warmup 1 --no-graph > gpurun_out/r02_config5_n1_eager.json 2> /dev/null
timeout 300 python tools/aux_timing.py > gpurun_out/r02_aux_timing.json 2> gpurun_out/r02_aux_timing.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n1.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "launches/step", d["gpu_launches_per_step"], "e2e", d["e2e"]["value"], "hr-e2e", d["e2e"]["hr_input_variant"]["value"])
print("kernels", {k: round(v["ms_per_view"] or v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
print("exact16", d["exact_tile16"]); print("fast_exp", d["fast_exp_blend"]); print("cpu", d["cpu_baseline"]); print("clocks", d["clocks"])
r = json.load(open("gpurun_out/r02_bench_reference.json")); print("reference", r["value"], r["cpu_baseline"]["cores"], r.get("gpu_reference_cuda"))
p = json.load(open("gpurun_out/r02_bench_perview.json")); print("per-view", p["value"], p["ms_per_step"], p["gpu_launches_per_step"])
c = json.load(open("gpurun_out/r02_config4.json")); print("config4", c["value"], c["ms_per_step"], c["roofline"]["encode_ms"], c["roofline"]["decode_ms"], c["roofline"]["frac"], c["e2e"])
for f in ("r02_config5_n1", "r02_config5_n1_eager"):
    c = json.load(open(f"gpurun_out/{f}.json")); print(f, c["value"], c["tracking_ms_per_iteration"], c["mapping_ms_per_iteration"], c["ae_ms_per_keyframe"], c.get("gpu_launches_per_iteration"))
PY
