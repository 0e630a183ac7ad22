set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_blend_bwd|k_blend2' -s 6 -c 2 -o gpurun_out/r2g_v2 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hr --no-graph > /dev/null 2> gpurun_out/r2g.err; tail -2 gpurun_out/r2g.err
OLS_BWD_V1=1 timeout 300 ncu --set full --clock-control none -k regex:'k_blend_bwd' -s 3 -c 1 -o gpurun_out/r2g_v1 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hr --no-graph > /dev/null 2>> gpurun_out/r2g.err
