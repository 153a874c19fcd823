#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py > gpurun_out/mgpu_check.log 2>&1; echo "check rc=$?"; tail -12 gpurun_out/mgpu_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_mgpu_$N.json 2> gpurun_out/bench_mgpu_$N.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_mgpu_$N.err; cat gpurun_out/bench_mgpu_$N.json
