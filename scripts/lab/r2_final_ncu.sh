#!/bin/bash
mkdir -p gpurun_out
timeout 330 ncu --set full --clock-control none -k regex:"filter_kernel|fused_filter_kernel|cumsum_stream_kernel|take_partition_kernel|take_window_gather_kernel|take_unpermute_kernel" -c 18 -o gpurun_out/r2final_prof -f \
    python scripts/prof_kernels.py 40000000 1 > gpurun_out/r2final_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2final_ncu.log
python scripts/ncu_summary.py gpurun_out/r2final_prof.ncu-rep gpurun_out/r2final_ncu_kernels.csv
rm -f gpurun_out/r2final_prof.ncu-rep
