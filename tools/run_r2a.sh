# round 2, GPU run A: parity of the batched path + first numbers (every step under its own timeout)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -25 > gpurun_out/r2a_tests.log; tail -8 gpurun_out/r2a_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -5 gpurun_out/r2a_bench.err
OLS_BLEND_V1=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-hr > gpurun_out/r2a_bench_v1.json 2> gpurun_out/r2a_bench_v1.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-hr --per-view > gpurun_out/r2a_bench_perview.json 2> gpurun_out/r2a_bench_perview.err
timeout 300 python bench.py --config 4 --steps 6 --warmup 3 > gpurun_out/r2a_config4.json 2> gpurun_out/r2a_config4.err; tail -3 gpurun_out/r2a_config4.err
python - <<'PY'
import json
for f in ("r2a_bench", "r2a_bench_v1", "r2a_bench_perview"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "eager", round(d["launch_mode"]["ms_per_step_eager"], 3),
              "e2e", d["e2e"] and round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), "launches/step", d.get("gpu_launches_per_step"))
        for k, v in d["kernels"].items():
            print("   ", k, round(v["ms_per_launch"], 4), v.get("ms_per_view") and round(v["ms_per_view"], 4), v["algorithmic_GBps"] and round(v["algorithmic_GBps"]))
        print("   fast_exp", d.get("fast_exp_blend"))
    except Exception as e:
        print(f, "FAILED", e)
try:
    d = json.load(open("gpurun_out/r2a_config4.json"))
    print("config4", d["value"], d["ms_per_step"], d["roofline"], d["two_stage"], d["parity"], d["e2e"])
except Exception as e:
    print("config4 FAILED", e)
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_' -s 66 -c 22 -o gpurun_out/r2a_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hr --no-graph > /dev/null 2> gpurun_out/r2a_full.err; tail -2 gpurun_out/r2a_full.err
