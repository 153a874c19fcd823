#!/bin/bash
# One call: radix-pass design-choice sweep (bit-field variants) -> GPU tests + bench with the best, ncu source-level
# captures of the best v3 variant and of one early-publish variant.  Logs in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== sweep"; timeout 900 python scripts/sort_sweep.py 2> gpurun_out/sweep.err | tee gpurun_out/sweep.log | cut -c1-400
BEST=$(cat gpurun_out/best_cfg 2>/dev/null || echo 0)
echo "best cfg = $BEST"
export MHB_SORT_CFG=$BEST
echo "== pytest -m gpu (cfg $BEST)"
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== bench (cfg $BEST)"
MHB_VERBOSE=1 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; echo rc=$?
cat gpurun_out/bench_r1d.json; grep "mhb\]" gpurun_out/bench_r1d.err | sort -u | head -4; tail -3 gpurun_out/bench_r1d.err
for C in $BEST ${NCU_EXTRA:-393}; do
  echo "== ncu full: radix pass cfg $C (2M reads)"
  MHB_SORT_CFG=$C timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_radix_pass -s 3 -c 1 -o gpurun_out/prof_radix_r1d_$C -f \
     python bench.py --reads 2000000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full_$C.log 2>&1; echo rc=$?
  ncu -i gpurun_out/prof_radix_r1d_$C.ncu-rep --page raw --csv > gpurun_out/prof_radix_r1d_${C}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_radix_r1d_$C.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_radix_r1d_${C}_src.csv 2>/dev/null
done
ls -la gpurun_out | head -40
