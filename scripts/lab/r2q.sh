#!/bin/bash
TAG=${1:-r2q}
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"sort_digit_scatter_kernel" -s 1 -c 1 -o gpurun_out/${TAG}_prof -f python scripts/lab/r2q_once.py 32000000 > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/${TAG}_prof.ncu-rep
