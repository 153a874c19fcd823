#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 80 python scripts/sort_sweep.py "$1" 2> gpurun_out/sweep.err > gpurun_out/sweep.log
python - <<PY
import json
for r in json.load(open('gpurun_out/sort_sweep.json')):
    if 'wr2_pass_ms' in r:
        print('cfg %5d (0x%04x) ok=%s wr2 %.2f ms  const %.2f  wr3 %.2f ms' % (r['cfg'], max(0, r['cfg'] - 256), r['ok'], sum(r['wr2_pass_ms']) / 7, min(r.get('wr2_const_digit_pass_ms', [0])), sum(r['wr3_pass_ms']) / 10))
    else:
        print(r)
PY
