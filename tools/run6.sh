python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"])
for k,v in d["kernels"].items(): print(k, round(v["ms_per_launch"],4), v["algorithmic_GBps"])
print(d["clocks"])
PY
tail -3 gpurun_out/bench_ours.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --tile 16 --backward-mode exact 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('tile16/exact value', d['value']); [print(k, round(v['ms_per_launch'],4)) for k,v in d['kernels'].items()]"
