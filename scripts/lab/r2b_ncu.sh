#!/bin/bash
TAG=${1:-r2b}
KREGEX=${2:-cumsum_stream_kernel}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s 2 -c 1 -o gpurun_out/${TAG}_prof -f \
    python scripts/lab/r2b_lab.py 100000000 1 > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${TAG}_ncu.log; ls -la gpurun_out/*.ncu-rep
