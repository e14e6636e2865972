#!/bin/bash
TAG=${1:-r2b}
KREGEX=${2:-cumsum_stream_kernel}
SKIP=${3:-2}
CNT=${4:-1}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s $SKIP -c $CNT -o gpurun_out/${TAG}_prof -f \
    python scripts/lab/r2b_lab.py 100000000 1 > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${TAG}_ncu.log; ls -la gpurun_out/*.ncu-rep
