mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-hr > gpurun_out/r2q_a.json 2> /dev/null
OLS_BWD_GSM=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-hr > gpurun_out/r2q_b.json 2> /dev/null
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-hr --tile 16 --backward-mode exact > gpurun_out/r2q_c.json 2> /dev/null
OLS_BWD_GSM=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-hr --tile 16 --backward-mode exact > gpurun_out/r2q_d.json 2> /dev/null
OLS_BWD_GSM=1 timeout 600 python -m pytest tests/test_backward_gpu.py tests/test_batch_gpu.py tests/test_disentangle_gpu.py -m gpu -q 2>&1 | tail -3
python - <<'PY'
import json
for f in "abcd":
    try:
        d = json.load(open(f"gpurun_out/r2q_{f}.json"))
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3),
              " ".join(f"{k}={v['ms_per_view'] and round(v['ms_per_view'],4)}" for k, v in d["kernels"].items()))
    except Exception as e:
        print(f, "FAILED", e)
PY
