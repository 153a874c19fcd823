#!/bin/bash
# k-list throughput at N=1 (config 4's k values; 5 M reads per k so that the widest records fit comfortably) and the
# count stage in rounds at bench scale (A13 at a capacity-shaped cap).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 21 29 39 59 79 99 119 141; do
  timeout 300 python bench.py --k $k --reads 5000000 --steps 2 --warmup 2 --e2e-steps 1 --no-cpu-baseline > gpurun_out/klist_k$k.json 2> gpurun_out/klist_k$k.err
  python - $k <<PY
import json, sys
k = sys.argv[1]
try:
    j = json.loads([l for l in open('gpurun_out/klist_k%s.json' % k) if l.startswith('{')][-1])
    r = j['roofline']
    print('k=%s: %.1f ms/step  %.3g edges/s  e2e %.3g  records %d B, %d passes, pass %.2f ms frac %.3f  stages %s' % (
        k, j['ms_per_step'], j['value'], j['e2e']['value'] or 0, r['algorithmic_bytes_per_launch'] // 2 // j['config']['n_edge_records'],
        len(r['per_pass_ms']), r['avg_launch_ms'], r['frac'], {a: round(b, 1) for a, b in j['stage_ms'].items()}))
except Exception as e:
    print('k=%s unreadable' % k, e, open('gpurun_out/klist_k%s.err' % k).read()[-300:])
PY
done
echo "== count stage in rounds at bench scale (10 M reads, cap = 1/8 of the records)"
timeout 600 python - <<PY
import sys, time, numpy as np
sys.path.insert(0, '.')
from megahit_b200 import lib, synth
n_reads, k, m = 10_000_000, 27, 2
b = synth.synth_reads(n_reads, 150, 5 * n_reads, 0.01, seed=1).reshape(-1)
one = lib.count_host(b, n_reads, k, m, want_mercy=True)
lib.set_round_limit(n_reads * (150 - k) // 8)
t0 = time.time(); g = lib.count_host(b, n_reads, k, m, want_mercy=True); t1 = time.time()
lib.set_round_limit(0)
print('rounds', g['n_rounds'], 'wall %.2f s' % (t1 - t0), 'ms', {a: round(v, 1) for a, v in g['ms'].items()},
      'identical to one pass:', bool(g['n_solid'] == one['n_solid'] and (g['edges'] == one['edges']).all() and (g['cand_ids'] == one['cand_ids']).all() and (g['counting'] == one['counting']).all()))
print('one pass ms', {a: round(v, 1) for a, v in one['ms'].items()})
PY
