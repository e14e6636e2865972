#!/bin/bash
# take: cache / prefetch-size hints on the random gather (100M-row table and the C4 8 GB table)
mkdir -p gpurun_out
for m in 0 1 2 3 4 5 6; do
  echo "== AG_TAKE_GATHER=$m"
  AG_TAKE_GATHER=$m timeout 300 python scripts/prof_kernels.py 2>&1 | grep take_
  AG_TAKE_GATHER=$m timeout 300 python scripts/c4_shard.py 1000000000 125000000 3 2>&1 | tail -1 | cut -c1-260
done 2>&1 | tee gpurun_out/take_gather_modes.txt
