set -x
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/r02_topo.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; tail -3 gpurun_out/r02_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 --no-hr > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; tail -3 gpurun_out/r02_bench_n4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --config 5 --gpus 8 --steps 8 --warmup 1 > gpurun_out/r02_config5_n8.json 2> gpurun_out/r02_config5_n8.err; tail -3 gpurun_out/r02_config5_n8.err
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.load(open(f"gpurun_out/r02_bench_n{n}.json"))
        e = d["e2e"]
        print(f"N{n} value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", e and round(e["value"], 1), e and round(e["ms_per_step"], 2),
              e and e.get("h2d_ceiling"), "hr-e2e", e and e.get("hr_input_variant") and round(e["hr_input_variant"]["value"], 1), "reduce_check", d["reduce_check"] and d["reduce_check"]["rel_l2_reduced_vs_single_rank_sum"])
    except Exception as ex:
        print(n, "FAILED", ex)
try:
    d = json.load(open("gpurun_out/r02_config5_n8.json"))
    print("config5 N8 fps", d["value"], "track ms/it", d["tracking_ms_per_iteration"], "map ms/it", d["mapping_ms_per_iteration"])
except Exception as ex:
    print("config5 N8 FAILED", ex)
PY
