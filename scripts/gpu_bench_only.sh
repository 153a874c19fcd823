#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 2 --e2e-steps ${E2E:-2} --no-cpu-baseline "$@" > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_iter.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_iter.json'))
print('value %.3e edges/s  ms/step %.1f' % (j['value'], j['ms_per_step']))
print('stage_ms', {k: round(v,2) for k,v in j['stage_ms'].items()})
r=j['roofline']; print('radix avg ms %.2f frac %.3f' % (r['avg_launch_ms'], r['frac'])); print('per pass', [round(x,2) for x in r['per_pass_ms']])
print('s2s pass', r['s2s_pass']); print('e2e', j['e2e'])
PY
