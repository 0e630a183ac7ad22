set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python tools/quick_fb.py 1000000 5 2>&1 | tail -10
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 30 --csv --log-file gpurun_out/launches_fb.csv python tools/quick_fb.py 1000000 2 exact > gpurun_out/ncu_fb.log 2>&1
OLS_SKIP_REF=1 ncu --set full --clock-control none --import-source on -k regex:'k_blend|k_sort_tiles_radix|k_preprocess|k_scatter|k_tile_offsets' -s 16 -c 8 -o gpurun_out/prof_r1a python tools/quick_fb.py 1000000 1 exact > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
