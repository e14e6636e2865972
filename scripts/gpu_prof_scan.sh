#!/bin/bash
TAG=${1:-ps}
mkdir -p gpurun_out
for v in 0 1; do
  AG_SCAN_TMA=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:cumsum -s 1 -c 1 -o gpurun_out/${TAG}_scan$v -f python scripts/prof_scan.py > gpurun_out/${TAG}_ncu$v.log 2>&1
  echo "ncu rc=$?"
  ncu -i gpurun_out/${TAG}_scan$v.ncu-rep --page details > gpurun_out/${TAG}_details$v.txt 2>/dev/null
  ncu -i gpurun_out/${TAG}_scan$v.ncu-rep --page source --csv > gpurun_out/${TAG}_source$v.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_scan$v.ncu-rep
done
ls -la gpurun_out/${TAG}_*
