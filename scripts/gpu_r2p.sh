#!/bin/bash
# read2sdbg (A12) on the GPU: new tests first, then the whole GPU suite (the emitter gained a label format), timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_r2s.py -m gpu -q --timeout 600 --maxfail=8 --tb=short > gpurun_out/r2p_pytest_r2s.txt 2>&1
tail -40 gpurun_out/r2p_pytest_r2s.txt
timeout 300 python scripts/r2s_time.py 2000000 ref > gpurun_out/r2p_r2s_time.jsonl 2> gpurun_out/r2p_r2s_time.err; cat gpurun_out/r2p_r2s_time.jsonl; tail -3 gpurun_out/r2p_r2s_time.err
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gpu_r2s.py -x > gpurun_out/r2p_pytest_gpu.txt 2>&1
tail -5 gpurun_out/r2p_pytest_gpu.txt
