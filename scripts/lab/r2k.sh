#!/bin/bash
TAG=${1:-r2k}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/lab/r2k_once.py 100000000 > gpurun_out/${TAG}_once.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/${TAG}_once.log
timeout 300 python scripts/sort_bench.py 100000000 2>&1 | tee gpurun_out/${TAG}_sort.txt
