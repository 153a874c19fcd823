#!/bin/bash
# ncu --set full on selected kernels of a small bench run.  usage: gpu_ncu.sh <tag> <kernel-regex> <skip> <count>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=$1; K=$2; S=$3; C=$4
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -o gpurun_out/prof_$TAG -f \
   python bench.py --reads 2000000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_$TAG.log 2>&1; echo rc=$?
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-300
