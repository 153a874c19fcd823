#!/bin/bash
# Round 2, GPU call C (1 GPU): hashed count stage - parity tests, bench in both modes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "hashed or large_synthetic or mercy_host or plan_partition or owner_answered" > gpurun_out/pytest_c.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_c.log
summ() { python - "$1" <<PY
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r = j["roofline"]
    print(sys.argv[1], "ms/step %.1f  value %.3g  e2e %.3g (%.1f ms)  pass %.2f ms frac %.3f  stages %s" % (
        j["ms_per_step"], j["value"], j["e2e"]["value"] or 0, j["e2e"]["ms_per_step"] or 0, r["avg_launch_ms"], r["frac"],
        {k: round(v, 1) for k, v in j["stage_ms"].items()}))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
echo "== bench hashed"
MHB_VERBOSE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --count-mode hashed --e2e-steps 0 > gpurun_out/bench_hashed.json 2> gpurun_out/bench_hashed.err; summ gpurun_out/bench_hashed.json; grep "mhb\]" gpurun_out/bench_hashed.err | sort | uniq | head
echo "== bench sort (skipped)"; if false; then
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --count-mode sort --e2e-steps 0 > gpurun_out/bench_sortmode.json 2> gpurun_out/bench_sortmode.err; summ gpurun_out/bench_sortmode.json
fi; echo "== ncu of the hash-count kernel (2 M reads)"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_hash_count -c 1 -o gpurun_out/r2c_hash_count python bench.py --steps 1 --warmup 1 --no-cpu-baseline --count-mode hashed --e2e-steps 0 --reads 2000000 > gpurun_out/ncu_hc.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep 2>/dev/null
