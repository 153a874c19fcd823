#!/bin/bash
# bench (real numbers) + ncu launch list + one full ncu capture of the radix pass.  Logs in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${1:-r1}
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo rc=$?
cat gpurun_out/bench_$R.json; tail -5 gpurun_out/bench_$R.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$R.json 2>&1; cat gpurun_out/bench_ref_$R.json
echo "== ncu launch list (2M reads)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
   python bench.py --reads 2000000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_list_$R.log 2>&1; echo rc=$?
tail -3 gpurun_out/ncu_list_$R.log
echo "== ncu full: radix pass (2M reads)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_radix_pass -s 3 -c 2 -o gpurun_out/prof_radix_$R -f \
   python bench.py --reads 2000000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full_$R.log 2>&1; echo rc=$?
tail -3 gpurun_out/ncu_full_$R.log
ls -la gpurun_out
