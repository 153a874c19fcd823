#!/bin/bash
# Round 2, GPU call B (2 GPUs): new single-GPU tests, multi-GPU parity at N=2, bench N=2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== new single-GPU tests"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "mercy_host or plan_partition or owner_answered or every_pass_variant or count_in_rounds or fused_build_matches" > gpurun_out/pytest_b.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_b.log
echo "== parity at N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py > gpurun_out/mgpu_check_n2.log 2>&1; echo "rc=$?"; tail -20 gpurun_out/mgpu_check_n2.log
echo "== bench N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; tail -5 gpurun_out/bench_n2.err
python - <<PY
import json
try:
    j = json.loads([l for l in open('gpurun_out/bench_n2.json') if l.startswith('{')][-1])
    print('N=2: %.1f ms/step value %.3g e2e %.3g parity %s launches %s' % (j['ms_per_step'], j['value'], j['e2e']['value'], j['parity']['ok'], j['gpu_launches']))
    print({k: round(v, 1) for k, v in j['stage_ms_max_over_ranks'].items()})
except Exception as e:
    print('bench_n2 unreadable', e)
PY
