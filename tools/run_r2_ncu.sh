set -x
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_' -s 42 -c 14 -o gpurun_out/r02_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hr --no-graph > /dev/null 2> gpurun_out/r02_full.err; tail -2 gpurun_out/r02_full.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'.' -s 0 -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-hr --no-graph > /dev/null 2> gpurun_out/r02_launches.err
