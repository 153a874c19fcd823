#!/bin/bash
# quick check: new tests + default bench with the mercy-edge stage inside the device step and the sampled index
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "odd_length or fused_build or relaxed" 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2o.json 2> gpurun_out/bench_r2o.err; tail -2 gpurun_out/bench_r2o.err
python - <<PY
import json
try:
    j = json.loads([l for l in open('gpurun_out/bench_r2o.json') if l.startswith('{')][-1]); r = j['roofline']
    print('ms/step %.1f value %.3g e2e %.3g (%.1f ms) pass %.2f ms frac %.3f launches %s cpu %.3g' % (j['ms_per_step'], j['value'], j['e2e']['value'], j['e2e']['ms_per_step'], r['avg_launch_ms'], r['frac'], j['gpu_launches'], j['cpu_baseline']['value']))
    print({k: round(v, 1) for k, v in j['stage_ms'].items()}, {k: round(v, 1) for k, v in j['e2e']['stages'].items() if isinstance(v, float)})
except Exception as e:
    print('unreadable', e)
PY
