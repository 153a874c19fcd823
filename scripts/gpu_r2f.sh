#!/bin/bash
# Round 2, GPU call F (1 GPU): full GPU suite, default bench (e2e + cpu baseline), launch list + ncu of the hash kernel at
# bench size, k-list sweep, rounds at bench scale.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rxXf > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench default"
timeout 600 python bench.py > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; tail -2 gpurun_out/bench_r2f.err
python - <<PY
import json
try:
    j = json.loads([l for l in open('gpurun_out/bench_r2f.json') if l.startswith('{')][-1]); r = j['roofline']
    print('ms/step %.1f value %.3g e2e %.3g (%.1f ms) pass %.2f ms frac %.3f launches %s cpu %.3g' % (j['ms_per_step'], j['value'], j['e2e']['value'], j['e2e']['ms_per_step'], r['avg_launch_ms'], r['frac'], j['gpu_launches'], j['cpu_baseline']['value']))
    print({k: round(v, 1) for k, v in j['stage_ms'].items()}, j['e2e']['stages'])
except Exception as e:
    print('unreadable', e)
PY
echo "== ncu launch list at bench size"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2f_10M.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_list.log 2>&1; echo rc=$?
echo "== ncu --set full of the hash kernel + one radix pass at bench size"
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:k_hash_count|k_radix_pass3' -c 2 --launch-skip 4 -o gpurun_out/r2f_hash_radix_10M \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full.log 2>&1; echo rc=$?
bash scripts/gpu_klist.sh
