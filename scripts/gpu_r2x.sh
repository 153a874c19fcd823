#!/bin/bash
# r2x: final validation of the round: whole GPU suite, smoke, default bench, iterate timing at 2 M reads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rxXf > gpurun_out/r2x_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2x_pytest_gpu.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== bench default"
timeout 600 python bench.py > gpurun_out/r2x_bench_default.json 2> gpurun_out/r2x_bench_default.err; tail -2 gpurun_out/r2x_bench_default.err
python - <<PY
import json
try:
    j = json.loads([l for l in open('gpurun_out/r2x_bench_default.json') if l.startswith('{')][-1]); r = j['roofline']
    print('ms/step %.1f value %.3g e2e %.3g (%.1f ms) pass %.2f ms frac %.3f traffic %s launches %s cpu %.3g' % (j['ms_per_step'], j['value'], j['e2e']['value'], j['e2e']['ms_per_step'], r['avg_launch_ms'], r['frac'], r['traffic'], j['gpu_launches'], j['cpu_baseline']['value']))
    print('  ', {k: round(v, 1) for k, v in j['stage_ms'].items()}, j.get('clocks'))
except Exception as e:
    print('unreadable', e)
PY
echo "== iterate at 2 M reads"
timeout 900 python scripts/iter_time.py 2000000 > gpurun_out/r2x_iter_time_2M.jsonl 2> gpurun_out/r2x_iter_time_2M.err; cat gpurun_out/r2x_iter_time_2M.jsonl; tail -3 gpurun_out/r2x_iter_time_2M.err
