set -x
mkdir -p gpurun_out
timeout 900 python bench.py --config 5 --steps 8 --warmup 1 > gpurun_out/r2d_config5.json 2> gpurun_out/r2d_config5.err; tail -8 gpurun_out/r2d_config5.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2d_config5.json"))
    print("config5 fps", d["value"], "ms/frame", d["ms_per_step"], "e2e", d["e2e"]["value"], d["breakdown_ms"], d["counts"],
          "track ms/it", d["tracking_ms_per_iteration"], "map ms/it", d["mapping_ms_per_iteration"], "ae ms/kf", d["ae_ms_per_keyframe"], d.get("gpu_launches_per_iteration"))
except Exception as e:
    print("config5 FAILED", e)
PY
