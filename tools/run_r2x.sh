OLS_AE_TRACE=1 timeout 120 python tools/ae_trace.py 2>&1 | grep "ae trace" | tail -1 | cut -c1-900
OLS_AE_L0_TF32=1 OLS_AE_TRACE=1 timeout 120 python tools/ae_trace.py 2>&1 | grep "ae trace" | tail -1 | cut -c1-900
