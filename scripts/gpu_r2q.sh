#!/bin/bash
# r2q: kmsort restructured (radix levels + one insertion thread per small range), phase trace of read2sdbg, and the
# pruned seq2sdbg extraction (aux flags) in the fused build and the bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_r2s.py -m gpu -q --timeout 600 --maxfail=8 --tb=short > gpurun_out/r2q_pytest_r2s.txt 2>&1
tail -4 gpurun_out/r2q_pytest_r2s.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_downstream.py -m gpu -q --timeout 600 -k "fused or build" --tb=short > gpurun_out/r2q_pytest_fused.txt 2>&1
tail -4 gpurun_out/r2q_pytest_fused.txt
timeout 300 python scripts/r2s_time.py 2000000 > gpurun_out/r2q_r2s_time_2M.jsonl 2> gpurun_out/r2q_r2s_time_2M.err; cat gpurun_out/r2q_r2s_time_2M.jsonl; grep "r2s\]" gpurun_out/r2q_r2s_time_2M.err | tail -60
timeout 400 python scripts/r2s_time.py 10000000 > gpurun_out/r2q_r2s_time_10M.jsonl 2> gpurun_out/r2q_r2s_time_10M.err; cat gpurun_out/r2q_r2s_time_10M.jsonl; grep "r2s\]" gpurun_out/r2q_r2s_time_10M.err | tail -60
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2q_bench_pruned.json 2> gpurun_out/r2q_bench_pruned.err; tail -2 gpurun_out/r2q_bench_pruned.err
MHB_S2S_NO_PRUNE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench_noprune.json 2> gpurun_out/r2q_bench_noprune.err; tail -2 gpurun_out/r2q_bench_noprune.err
python - <<PY
import json
for f in ("pruned", "noprune"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/r2q_bench_{f}.json") if l.startswith('{')][-1]); r = j['roofline']
        print(f, 'ms/step %.1f value %.3g e2e %.3g (%.1f ms) pass %.2f ms frac %.3f launches %s' % (j['ms_per_step'], j['value'], j['e2e']['value'], j['e2e']['ms_per_step'], r['avg_launch_ms'], r['frac'], j['gpu_launches']))
        print({k: round(v, 1) for k, v in j['stage_ms'].items()}, j['config'].get('n_sdbg_sort_items'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
