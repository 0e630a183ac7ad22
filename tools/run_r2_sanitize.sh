set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/r02_memcheck.log 2>&1; tail -3 gpurun_out/r02_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 6 python tools/sanitize.py > gpurun_out/r02_racecheck.log 2>&1; grep -E "hazard|Race|SUMMARY|sanitize pass" gpurun_out/r02_racecheck.log | head -12
timeout 300 python tools/aux_timing.py > gpurun_out/r02_aux_timing.json 2> gpurun_out/r02_aux_timing.err; tail -2 gpurun_out/r02_aux_timing.err; python -c "
import json; d=json.load(open('gpurun_out/r02_aux_timing.json')); [print(k, v['ms'], v['torch_reference_ms'], v['speedup']) for k,v in d.items()]"
