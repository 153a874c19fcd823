#!/bin/bash
# Last call of round 1: sweep of the early-scan / late-prefetch variants, then the full GPU test suite and the bench with
# the best one.  Logs in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== sweep"; timeout 240 python scripts/sort_sweep.py 2> gpurun_out/sweep.err > gpurun_out/sweep.log
python - <<PY
import json
for r in json.load(open('gpurun_out/sort_sweep.json')):
    if 'wr2_pass_ms' in r:
        print('cfg %4d (0x%04x) ok=%s wr2 %.2f ms  const %.2f  wr3 %.2f ms' % (r['cfg'], max(0, r['cfg'] - 256), r['ok'], sum(r['wr2_pass_ms']) / 7, min(r.get('wr2_const_digit_pass_ms', [0])), sum(r['wr3_pass_ms']) / 10))
    else:
        print(r)
PY
BEST=$(cat gpurun_out/best_cfg 2>/dev/null || echo 384)
echo "best cfg = $BEST"
export MHB_SORT_CFG=$BEST
echo "== pytest -m gpu (cfg $BEST)"
timeout 300 python -m pytest tests -m gpu -q --timeout 200 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== bench (cfg $BEST)"
MHB_VERBOSE=1 timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; echo rc=$?
cat gpurun_out/bench_r1f.json; grep "mhb\]" gpurun_out/bench_r1f.err | sort -u | head -4; tail -3 gpurun_out/bench_r1f.err
