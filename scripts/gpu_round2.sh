#!/bin/bash
# Full GPU round: parity suite + bench (both arms) + ncu launch list + ncu full captures of every hot kernel.
# Usage (repo root, on the GPU box):  bash scripts/gpu_round2.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/${TAG}_box.txt
echo "nproc=$(nproc)" >> gpurun_out/${TAG}_box.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/${TAG}_box.txt
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python scripts/prof_kernels.py > gpurun_out/${TAG}_kernels.txt 2>&1; cat gpurun_out/${TAG}_kernels.txt
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 5 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
# launch list of the bench command (cold-cache, serialised: compare SHARES, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
# the dominant kernel of the bench step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:binary_spans_kernel -s 3 -c 2 -o gpurun_out/${TAG}_prof_add -f \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/${TAG}_ncu_add.log 2>&1
echo "ncu add rc=$?"
# every other hot kernel, one capture each
timeout 900 ncu --set full --clock-control none -k regex:'sum_kernel|compare_kernel|filter_kernel|take_kernel|checked_tile_kernel|cast_vec_kernel|minmax_kernel|cumsum_kernel' -c 48 -o gpurun_out/${TAG}_prof_kernels -f \
    python scripts/prof_kernels.py 100000000 1 > gpurun_out/${TAG}_ncu_kernels.log 2>&1
echo "ncu kernels rc=$?"
# gpurun brings back at most 64 MiB: keep the compact per-launch CSVs, drop the raw reports
python scripts/ncu_summary.py gpurun_out/${TAG}_prof_add.ncu-rep gpurun_out/${TAG}_ncu_add.csv
python scripts/ncu_summary.py gpurun_out/${TAG}_prof_kernels.ncu-rep gpurun_out/${TAG}_ncu_kernels.csv
ls -la gpurun_out/*.ncu-rep; rm -f gpurun_out/*.ncu-rep
