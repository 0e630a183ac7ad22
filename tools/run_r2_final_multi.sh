set -x
mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; grep "balance rank 0" gpurun_out/r02_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --config 5 --gpus $N --steps 8 --warmup 1 > gpurun_out/r02_config5_n$N.json 2> gpurun_out/r02_config5_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.load(open(f"gpurun_out/r02_bench_n{n}.json")); e = d["e2e"]
print(f"N{n} value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(e["value"], 1), round(e["ms_per_step"], 2), e["h2d_ceiling"], "floor", round(e["transfer_floor_ms_per_step"], 2),
      "hr-e2e", e.get("hr_input_variant") and round(e["hr_input_variant"]["value"], 1), "reduce_check", d["reduce_check"]["rel_l2_reduced_vs_single_rank_sum"], d["rank_render_ms_after_balancing"])
c = json.load(open(f"gpurun_out/r02_config5_n{n}.json")); print("config5", c["value"], c["tracking_ms_per_iteration"], c["mapping_ms_per_iteration"], c["ae_ms_per_keyframe"])
PY
