#!/bin/bash
TAG=${1:-ps1}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cumsum_kernel -s 1 -c 1 -o gpurun_out/${TAG}_scan -f python scripts/prof_scan.py > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}_scan.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_scan.ncu-rep --page details > gpurun_out/${TAG}_details.txt 2>/dev/null
ncu -i gpurun_out/${TAG}_scan.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
ls -la gpurun_out/${TAG}_*; rm -f gpurun_out/${TAG}_scan.ncu-rep
