#!/bin/bash
# Round 2, GPU call I (8 GPUs): parity + C++ CLI test at N, then bench --gpus 8 at config-3 size (100 M reads over 8 GPUs
# = 12.5 M reads per GPU) and at the weak-scaling size of the other runs (10 M per GPU).
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== C++ multi-GPU CLI at min(N,4) GPUs"
timeout 600 python -m pytest tests/test_gpu_downstream.py -m gpu -q --timeout 600 -k "multi_gpu" > gpurun_out/pytest_i.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_i.log
echo "== parity at N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/mgpu_check.py > gpurun_out/mgpu_check_n$N.log 2>&1; echo "rc=$?"; grep -E "MGPU PARITY|identical|MISMATCH|False" gpurun_out/mgpu_check_n$N.log | tail -6
for R in 10000000 12500000; do
echo "== bench N=$N, $R reads per GPU"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 3 --warmup 2 --reads $R > gpurun_out/bench_n${N}_$R.json 2> gpurun_out/bench_n${N}_$R.err; echo "rc=$?"; tail -2 gpurun_out/bench_n${N}_$R.err
python - $N $R <<PY
import json, sys
N, R = sys.argv[1], sys.argv[2]
try:
    j = json.loads([l for l in open('gpurun_out/bench_n%s_%s.json' % (N, R)) if l.startswith('{')][-1])
    print('N=%s reads/GPU %s: %.1f ms/step value %.3g e2e %.3g parity %s launches %s' % (N, R, j['ms_per_step'], j['value'], j['e2e']['value'], j['parity']['ok'], j['gpu_launches']))
    print({k: round(v, 1) for k, v in j['stage_ms_max_over_ranks'].items()})
except Exception as e:
    print('bench unreadable', e)
PY
done
