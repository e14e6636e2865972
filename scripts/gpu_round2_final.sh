#!/bin/bash
# Round-2 closing GPU run: full parity suite, both bench arms, ncu launch list of the bench command, ncu --set full of
# the kernels that changed last (sort / hash).   bash scripts/gpu_round2_final.sh <tag>
TAG=${1:-r2z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/${TAG}_box.txt
echo "nproc=$(nproc)" >> gpurun_out/${TAG}_box.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/${TAG}_box.txt
timeout 420 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
timeout 420 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 200 python bench.py --impl reference --steps 5 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
echo "ref rc=$?"
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
timeout 200 ncu --set full --clock-control none -k regex:"sort_prep_kernel|sort_digit_hist_kernel|sort_digit_scatter_kernel|sort_scan_bins_kernel|hash_insert_kernel|unique_mark_kernel|is_in_kernel" -c 24 -o gpurun_out/${TAG}_prof_sorthash -f \
    python scripts/lab/r2k_once.py 32000000 > gpurun_out/${TAG}_ncu_sorthash.log 2>&1
echo "ncu sorthash rc=$?"
python scripts/ncu_summary.py gpurun_out/${TAG}_prof_sorthash.ncu-rep gpurun_out/${TAG}_ncu_sorthash.csv
rm -f gpurun_out/${TAG}_prof_sorthash.ncu-rep
