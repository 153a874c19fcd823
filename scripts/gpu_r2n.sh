#!/bin/bash
# Round 2, GPU call N (1 GPU): final validation of the default configuration + ncu evidence for the new kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rxXf > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== bench default"
timeout 600 python bench.py > gpurun_out/bench_r2n.json 2> gpurun_out/bench_r2n.err; tail -2 gpurun_out/bench_r2n.err
python - <<PY
import json
try:
    j = json.loads([l for l in open('gpurun_out/bench_r2n.json') if l.startswith('{')][-1]); r = j['roofline']
    print('ms/step %.1f value %.3g e2e %.3g (%.1f ms) pass %.2f ms frac %.3f launches %s cpu %.3g' % (j['ms_per_step'], j['value'], j['e2e']['value'], j['e2e']['ms_per_step'], r['avg_launch_ms'], r['frac'], j['gpu_launches'], j['cpu_baseline']['value']))
    print('per pass frac', [round(x, 3) for x in r['per_pass_frac']], 's2s', round(r['s2s_pass']['frac'], 3), j['config'].get('host_affinity'))
    print({k: round(v, 1) for k, v in j['stage_ms'].items()}, {k: round(v, 1) for k, v in j['e2e']['stages'].items() if isinstance(v, float)})
except Exception as e:
    print('unreadable', e)
PY
echo "== bench reference arm"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400
echo "== ncu launch list at bench size"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2n_10M.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_list.log 2>&1; echo rc=$?
echo "== ncu --set full: unstable pass, stable pass, hash kernel at bench size"
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:k_hash_count|k_part_unstable' -c 2 -o gpurun_out/r2n_hash_part_10M \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full.log 2>&1; echo rc=$?
