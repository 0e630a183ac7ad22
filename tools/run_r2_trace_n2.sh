mkdir -p gpurun_out
OLS_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-hr > gpurun_out/r2n_bench_n2.json 2> gpurun_out/r2n_bench_n2.err; grep "phase trace\|balance rank" gpurun_out/r2n_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_n2.json')); print(d['value'], d['ms_per_step'], d['rank_render_ms_after_balancing'])"
