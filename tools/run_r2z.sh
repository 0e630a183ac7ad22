timeout 300 python -m pytest tests/test_losses.py tests/test_tracking_gpu.py -m gpu -q 2>&1 | tail -3
timeout 120 python tools/loss_probe.py 2>&1 | tail -3
timeout 600 python bench.py --config 5 --steps 8 --warmup 1 > gpurun_out/r2z_config5.json 2> /dev/null; python -c "
import json; c=json.load(open('gpurun_out/r2z_config5.json')); print('config5', c['value'], c['tracking_ms_per_iteration'], c['mapping_ms_per_iteration'], c['ae_ms_per_keyframe'])"
