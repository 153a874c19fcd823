#!/bin/bash
# r2r: kmsort on shared memory (CTA per bucket on tags, warp per staged range), mercy early-out
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_r2s.py -m gpu -q --timeout 600 --maxfail=8 --tb=short > gpurun_out/r2r_pytest_r2s.txt 2>&1
tail -12 gpurun_out/r2r_pytest_r2s.txt
timeout 300 python scripts/r2s_time.py 2000000 > gpurun_out/r2r_r2s_time_2M.jsonl 2> gpurun_out/r2r_r2s_time_2M.err; cat gpurun_out/r2r_r2s_time_2M.jsonl; grep "r2s\]" gpurun_out/r2r_r2s_time_2M.err | head -20
timeout 400 python scripts/r2s_time.py 10000000 > gpurun_out/r2r_r2s_time_10M.jsonl 2> gpurun_out/r2r_r2s_time_10M.err; cat gpurun_out/r2r_r2s_time_10M.jsonl; grep "r2s\]" gpurun_out/r2r_r2s_time_10M.err | head -20
