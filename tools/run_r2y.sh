timeout 120 python tools/loss_probe.py 2>&1 | tail -3
timeout 200 ncu --set full --clock-control none -k regex:k_mapping_loss -s 12 -c 2 -o gpurun_out/r2y_loss python tools/loss_probe.py > /dev/null 2>&1
