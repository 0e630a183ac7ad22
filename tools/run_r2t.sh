mkdir -p gpurun_out
for c in 132 140; do
OLS_AE_MAX_CTAS=$c OLS_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-hr > gpurun_out/r2t_$c.json 2> gpurun_out/r2t_$c.err; grep "phase trace rank 0" gpurun_out/r2t_$c.err
python -c "
import json; d=json.load(open('gpurun_out/r2t_$c.json')); print($c, d['value'], d['ms_per_step'])"
done
