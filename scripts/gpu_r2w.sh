#!/bin/bash
# r2w: iterate (N2) on the GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_iter.py -m gpu -q -s --timeout 800 --maxfail=8 --tb=short > gpurun_out/r2w_pytest_iter.txt 2>&1
grep -v "^$" gpurun_out/r2w_pytest_iter.txt | tail -25
