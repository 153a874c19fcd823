#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for c in 0 1 2 3; do
  MHB_SORT_CFG=$c MHB_VERBOSE=1 timeout 600 python bench.py --steps 2 --warmup 2 --e2e-steps 0 --no-cpu-baseline > gpurun_out/sweep_$c.json 2> gpurun_out/sweep_$c.err
  grep "mhb\]" gpurun_out/sweep_$c.err | sort -u
  python - <<PY
import json
j=json.load(open('gpurun_out/sweep_$c.json'))
r=j['roofline']
print('cfg $c: ms/step %.1f  count-pass avg %.2f ms (%.3f)  s2s-pass %.2f ms (%.3f)  stages %s' % (j['ms_per_step'], r['avg_launch_ms'], r['frac'], r['s2s_pass']['avg_launch_ms'], r['s2s_pass']['frac'], {k: round(v,1) for k,v in j['stage_ms'].items()}))
PY
done
