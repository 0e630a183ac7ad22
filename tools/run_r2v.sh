mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_ae.py -m gpu -q -x --timeout 120 2>&1 | tail -8
timeout 120 python tools/ae_errors.py 2>&1 | grep -v fp32 | tail -4
OLS_AE_L0_TF32=1 timeout 120 python tools/ae_errors.py 2>&1 | grep -v fp32 | tail -4
