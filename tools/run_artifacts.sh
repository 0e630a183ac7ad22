# produces the round's measured artefacts under gpurun_out/ (copied into profiles/ afterwards); every step under its own timeout
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r01_bench_n1.json 2> gpurun_out/r01_bench_n1.err; tail -3 gpurun_out/r01_bench_n1.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference.json 2> gpurun_out/r01_bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 160 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-hr --no-graph > /dev/null 2> gpurun_out/r01_launches.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_blend|k_ae_chain' -s 30 -c 3 -o gpurun_out/r01_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hr --no-graph > /dev/null 2> gpurun_out/r01_full.err
timeout 200 python tools/hr_debug.py 2>&1 | tail -3
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/memcheck.log 2>&1; tail -1 gpurun_out/memcheck.log
timeout 300 compute-sanitizer --tool racecheck --print-limit 4 python tools/sanitize.py > gpurun_out/racecheck.log 2>&1
grep -E "hazard|Race|at .*\+0x|SUMMARY" gpurun_out/racecheck.log | head -16
OLS_HR_NO_SPLIT=1 timeout 300 compute-sanitizer --tool racecheck python tools/sanitize.py 2>&1 | tail -1
nproc; lscpu | grep "Model name"
