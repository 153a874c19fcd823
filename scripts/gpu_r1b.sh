#!/bin/bash
# One call: radix-pass variant sweep -> full GPU test suite, bench, ncu launch list and one full capture with the
# best variant.  Logs in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
echo "== sweep"; timeout 900 python scripts/sort_sweep.py 2> gpurun_out/sweep.err | tee gpurun_out/sweep.log
BEST=$(cat gpurun_out/best_cfg 2>/dev/null || echo 0)
echo "best cfg = $BEST"
export MHB_SORT_CFG=$BEST
echo "== pytest -m gpu (cfg $BEST)"
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== bench (cfg $BEST)"
MHB_VERBOSE=1 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo rc=$?
cat gpurun_out/bench_r1b.json; grep "mhb\]" gpurun_out/bench_r1b.err | sort -u | head; tail -3 gpurun_out/bench_r1b.err
echo "== bench cfg 0 (v2 baseline, same box)"
MHB_SORT_CFG=0 timeout 300 python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_r1b_cfg0.json 2> gpurun_out/bench_r1b_cfg0.err; echo rc=$?
python - <<PY
import json
for f in ("bench_r1b", "bench_r1b_cfg0"):
    try:
        j = json.load(open("gpurun_out/%s.json" % f)); r = j["roofline"]
        print(f, "ms/step %.1f value %.3g  count-pass %.2f ms frac %.3f  s2s-pass %.2f ms frac %.3f  stages %s e2e %s" % (
            j["ms_per_step"], j["value"], r["avg_launch_ms"], r["frac"], r["s2s_pass"]["avg_launch_ms"], r["s2s_pass"]["frac"],
            {k: round(v, 1) for k, v in j["stage_ms"].items()}, j["e2e"]["value"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
echo "== ncu launch list (2M reads)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv \
   python bench.py --reads 2000000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_list.log 2>&1; echo rc=$?
echo "== ncu full: radix pass (2M reads)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_radix_pass -s 3 -c 2 -o gpurun_out/prof_radix_r1b -f \
   python bench.py --reads 2000000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full.log 2>&1; echo rc=$?
ncu -i gpurun_out/prof_radix_r1b.ncu-rep --page raw --csv > gpurun_out/prof_radix_r1b_raw.csv 2>/dev/null
ls -la gpurun_out | head -40
