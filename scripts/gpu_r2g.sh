#!/bin/bash
# Round 2, GPU call G (2 GPUs): C++ multi-GPU CLI test, Python multi-GPU parity + bench at N=2 with the finer marks.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== C++ multi-GPU CLI"
timeout 600 python -m pytest tests/test_gpu_downstream.py -m gpu -q --timeout 600 -k "multi_gpu" -x > gpurun_out/pytest_g.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_g.log
echo "== timing of the C++ driver on 2 M reads"
python - <<PY
import os, sys, time, subprocess, tempfile
sys.path.insert(0, '.')
from megahit_b200 import formats as F, synth
d = tempfile.mkdtemp()
n = 2_000_000
b = synth.synth_reads(n, 150, 5 * n, 0.01, seed=5)
F.write_lib(d + '/r', b, n, n * 150, 150)
for g in (1, 2):
    t0 = time.time()
    r = subprocess.run(['megahit_b200/bin/megahit_core', 'count', '-k', '27', '-m', '2', '--host_mem', '1e10', '--mem_flag', '1', '--output_prefix', d + '/o%d' % g, '--num_cpu_threads', '8', '--read_lib_file', d + '/r', '--gpus', str(g)], capture_output=True, text=True)
    print('gpus', g, 'rc', r.returncode, 'wall %.2f s' % (time.time() - t0), r.stderr.strip().splitlines()[-1][-120:])
PY
bash scripts/gpu_r2d.sh 2
