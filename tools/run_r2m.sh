set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err; tail -5 gpurun_out/r2m_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --config 5 --gpus 2 --steps 8 --warmup 1 > gpurun_out/r2m_config5_n2.json 2> gpurun_out/r2m_config5_n2.err; tail -5 gpurun_out/r2m_config5_n2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2m_bench_n2.json"))
    print("N2 value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", d["e2e"] and round(d["e2e"]["value"], 1), d["e2e"] and d["e2e"]["ms_per_step"], "reduce_check", d["reduce_check"])
except Exception as e:
    print("N2 FAILED", e)
try:
    d = json.load(open("gpurun_out/r2m_config5_n2.json"))
    print("config5 N2 fps", d["value"], "track ms/it", d["tracking_ms_per_iteration"], "map ms/it", d["mapping_ms_per_iteration"])
except Exception as e:
    print("config5 N2 FAILED", e)
PY
