#!/bin/bash
# r2y: last sanity run of the round after the host-side clean-ups (result guard, RAII events): read2sdbg + iterate suites
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_r2s.py tests/test_gpu_iter.py tests/test_gpu_parity.py -m gpu -q --timeout 500 -k "not at_300k" --tb=short > gpurun_out/r2y_pytest.txt 2>&1
tail -5 gpurun_out/r2y_pytest.txt
