#!/bin/bash
TAG=${1:-r2m}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sort_digit_scatter_kernel|is_in_kernel|sort_digit_hist_kernel|sort_prep_kernel|unique_mark_kernel" -c 7 -o gpurun_out/${TAG}_prof -f python scripts/lab/r2k_once.py 32000000 > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${TAG}_ncu.log; ls -la gpurun_out/${TAG}_prof.ncu-rep
