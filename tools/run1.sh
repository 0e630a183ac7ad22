set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python tools/quick_fwd.py 1000000 2>&1 | tail -8
ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file gpurun_out/launches_fwd.csv python tools/quick_fwd.py 1000000 > gpurun_out/ncu_fwd.log 2>&1
tail -3 gpurun_out/ncu_fwd.log
