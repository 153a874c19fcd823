#!/bin/bash
# r2v: read2sdbg with shared stage buffers, fall-back paths of the kmsort emulation under test
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_r2s.py -m gpu -q --timeout 600 --maxfail=8 --tb=short > gpurun_out/r2v_pytest_r2s.txt 2>&1
tail -8 gpurun_out/r2v_pytest_r2s.txt
timeout 400 python scripts/r2s_time.py 10000000 > gpurun_out/r2v_r2s_time_10M.jsonl 2> gpurun_out/r2v_r2s_time_10M.err; cat gpurun_out/r2v_r2s_time_10M.jsonl; grep "r2s\]" gpurun_out/r2v_r2s_time_10M.err | head -26
