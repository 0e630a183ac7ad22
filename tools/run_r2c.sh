set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2c_tests.log 2>&1; tail -12 gpurun_out/r2c_tests.log
timeout 600 python bench.py --config 5 --steps 8 --warmup 1 > gpurun_out/r2c_config5.json 2> gpurun_out/r2c_config5.err; tail -5 gpurun_out/r2c_config5.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2c_config5.json"))
    print("config5 fps", d["value"], "ms/frame", d["ms_per_step"], "e2e", d["e2e"]["value"], d["breakdown_ms"], d["counts"],
          "track ms/it", d["tracking_ms_per_iteration"], "map ms/it", d["mapping_ms_per_iteration"], "ae ms/kf", d["ae_ms_per_keyframe"])
except Exception as e:
    print("config5 FAILED", e)
PY
