#!/bin/bash
# sort tests + probe + short bench (no e2e)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "sort or medium or large" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_quick.log
timeout 300 python scripts/sort_probe.py 246e6 2>&1 | tail -5
timeout 600 python bench.py --steps 2 --warmup 2 --e2e-steps 0 --no-cpu-baseline > gpurun_out/quick.json 2> gpurun_out/quick.err; tail -2 gpurun_out/quick.err
python - <<PY
import json
j=json.load(open('gpurun_out/quick.json'))
r=j['roofline']
print('ms/step %.1f  count-pass avg %.2f ms (%.3f)  s2s-pass %.2f ms (%.3f)  stages %s' % (j['ms_per_step'], r['avg_launch_ms'], r['frac'], r['s2s_pass']['avg_launch_ms'], r['s2s_pass']['frac'], {k: round(v,1) for k,v in j['stage_ms'].items()}))
PY
