#!/bin/bash
# Per-tile timeline of the radix pass (diagnostic build) for the default variant and the non-early base.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MHB_LIB=$PWD/megahit_b200/libmhb_timeline.so
for C in 0x080 0x000; do
  timeout 45 python scripts/sort_timeline.py 1.23e9 $((256 + C)) 2 2>&1 | tee gpurun_out/timeline_$C.txt
done
