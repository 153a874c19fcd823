#!/bin/bash
# One call: radix-pass sweep #4 (early publish x geometry / occupancy / look-back) -> GPU tests + bench with the best,
# ncu captures (CSV exports only: the .ncu-rep files are too big to travel back).  Logs in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== sweep"; timeout 900 python scripts/sort_sweep.py 2> gpurun_out/sweep.err > gpurun_out/sweep.log
python - <<PY
import json
for r in json.load(open('gpurun_out/sort_sweep.json')):
    if 'wr2_pass_ms' in r:
        print('cfg %4d (0x%03x) ok=%s wr2 %.2f ms  const %.2f  wr3 %.2f ms' % (r['cfg'], max(0, r['cfg'] - 256), r['ok'], sum(r['wr2_pass_ms']) / 7, min(r.get('wr2_const_digit_pass_ms', [0])), sum(r['wr3_pass_ms']) / 10))
    else:
        print(r)
PY
BEST=$(cat gpurun_out/best_cfg 2>/dev/null || echo 384)
echo "best cfg = $BEST"
export MHB_SORT_CFG=$BEST
echo "== pytest -m gpu (cfg $BEST)"
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== bench (cfg $BEST)"
MHB_VERBOSE=1 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err; echo rc=$?
cat gpurun_out/bench_r1e.json; grep "mhb\]" gpurun_out/bench_r1e.err | sort -u | head -4; tail -3 gpurun_out/bench_r1e.err
echo "== ncu full, source level: radix pass cfg $BEST (2M reads)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_radix_pass -s 3 -c 1 -o gpurun_out/prof_radix_r1e -f \
   python bench.py --reads 2000000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full.log 2>&1; echo rc=$?
ncu -i gpurun_out/prof_radix_r1e.ncu-rep --page raw --csv > gpurun_out/prof_radix_r1e_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_radix_r1e.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_radix_r1e_src.csv 2>/dev/null
rm -f gpurun_out/prof_radix_r1e.ncu-rep
echo "== ncu full at bench size (10M reads): DRAM traffic per launch of the radix pass"
timeout 600 ncu --set full --clock-control none -k regex:k_radix_pass -s 3 -c 2 -o gpurun_out/prof_radix_r1e_10M -f \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full_10M.log 2>&1; echo rc=$?
ncu -i gpurun_out/prof_radix_r1e_10M.ncu-rep --page raw --csv > gpurun_out/prof_radix_r1e_10M_raw.csv 2>/dev/null
rm -f gpurun_out/prof_radix_r1e_10M.ncu-rep
echo "== ncu launch list (2M reads)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1e.csv \
   python bench.py --reads 2000000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_list.log 2>&1; echo rc=$?
ls -la gpurun_out | head -40
du -sh gpurun_out
