python -m pytest tests -m gpu -x -q 2>&1 | tail -5
OLS_SKIP_REF=1 python tools/quick_fb.py 1000000 5 exact 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 10 --csv --log-file gpurun_out/launches_fb2.csv python tools/quick_fb.py 1000000 1 exact > gpurun_out/ncu_fb2.log 2>&1
