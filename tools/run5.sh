python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 3000 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
nproc; lscpu | grep "Model name"
