#!/bin/bash
# Round 2, GPU call D (N GPUs, N = $1): multi-GPU parity incl. per-rank files + reference assemble, then bench --gpus N.
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== parity at N=$N (fused path, files, reference assemble on the $N-file SdBG)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/mgpu_check.py > gpurun_out/mgpu_check_n$N.log 2>&1; echo "rc=$?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/mgpu_check_n$N.log | tail -16
echo "== bench N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; tail -3 gpurun_out/bench_n$N.err
python - $N <<PY
import json, sys
N = sys.argv[1]
try:
    j = json.loads([l for l in open('gpurun_out/bench_n%s.json' % N) if l.startswith('{')][-1])
    print('N=%s: %.1f ms/step value %.3g e2e %.3g parity %s launches %s' % (N, j['ms_per_step'], j['value'], j['e2e']['value'], j['parity']['ok'], j['gpu_launches']))
    print({k: round(v, 1) for k, v in j['stage_ms_max_over_ranks'].items()})
    print('owned', j['config']['records_owned_per_rank'])
except Exception as e:
    print('bench unreadable', e)
PY
