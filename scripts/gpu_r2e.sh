#!/bin/bash
# hash-count tuning: kernel geometry x slice size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== hashed tests for each geometry"
for G in A B C; do
  MHB_HC_GEOM=$G timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "hashed" 2>&1 | tail -1
done
for GS in "A 0" "B 0" "B 4500" "B 7500" "C 0" "C 9000" "C 15000"; do
  set -- $GS
  MHB_VERBOSE=1 MHB_HC_GEOM=$1 MHB_HC_SLICE=$2 timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --count-mode hashed --e2e-steps 0 > gpurun_out/bench_hc_$1_$2.json 2> gpurun_out/bench_hc_$1_$2.err
  python - $1 $2 <<PY
import json, sys
try:
    j = json.loads([l for l in open('gpurun_out/bench_hc_%s_%s.json' % (sys.argv[1], sys.argv[2])) if l.startswith('{')][-1])
    print('geom', sys.argv[1], 'slice', sys.argv[2], 'ms/step %.1f' % j['ms_per_step'], {k: round(v, 1) for k, v in j['stage_ms'].items()})
except Exception as e:
    print('geom', sys.argv[1], sys.argv[2], 'unreadable', e)
PY
  grep "hash count" gpurun_out/bench_hc_$1_$2.err | head -1
done
