#!/bin/bash
# r2t: the fused build falling into rounds (A13), and the k-list with the seq2sdbg item pruning
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "fused" --tb=short > gpurun_out/r2t_pytest_fused.txt 2>&1
tail -6 gpurun_out/r2t_pytest_fused.txt
for k in 21 29 39 59 79 99 119 141; do
  timeout 300 python bench.py --k $k --reads 5000000 --steps 2 --warmup 2 --e2e-steps 1 --no-cpu-baseline > gpurun_out/klist_k$k.json 2> gpurun_out/klist_k$k.err
  python - $k <<PY
import json, sys
k = sys.argv[1]
try:
    j = json.loads([l for l in open('gpurun_out/klist_k%s.json' % k) if l.startswith('{')][-1])
    r = j['roofline']
    print('k=%s: %.1f ms/step  %.3g edges/s  e2e %.3g  records %d B, %d passes, pass %.2f ms frac %.3f  s2s items %d pass frac %.3f  stages %s' % (
        k, j['ms_per_step'], j['value'], j['e2e']['value'] or 0, r['algorithmic_bytes_per_launch'] // 2 // j['config']['n_edge_records'],
        len(r['per_pass_ms']), r['avg_launch_ms'], r['frac'], j['config']['n_sdbg_sort_items'], r['s2s_pass']['frac'], {a: round(b, 1) for a, b in j['stage_ms'].items()}))
except Exception as e:
    print('k=%s unreadable' % k, e, open('gpurun_out/klist_k%s.err' % k).read()[-300:])
PY
done
