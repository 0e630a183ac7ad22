# produces the round's measured artefacts under gpurun_out/ (copied into profiles/ afterwards)
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r01_bench_n1.json 2> gpurun_out/r01_bench_n1.err; tail -3 gpurun_out/r01_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference.json 2> gpurun_out/r01_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 160 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-hr --no-graph > /dev/null 2> gpurun_out/r01_launches.err
ncu --set full --clock-control none --import-source on -k regex:'k_blend|k_ae_chain' -s 30 -c 3 -o gpurun_out/r01_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hr --no-graph > /dev/null 2> gpurun_out/r01_full.err
nproc; lscpu | grep "Model name"
