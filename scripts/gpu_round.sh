#!/bin/bash
# One gpurun call: GPU parity suite + bench + ncu launch list (+ optional full capture).
# Usage (from the repo root on the GPU box):  bash scripts/gpu_round.sh [tag] [ncu-kernel-regex]
TAG=${1:-r1}
KRE=${2:-}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/${TAG}_box.txt
echo "nproc=$(nproc)" >> gpurun_out/${TAG}_box.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/${TAG}_box.txt
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 5 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_ref.json
# launch list of the same command (cold-cache, serialised: compare SHARES, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
if [ -n "$KRE" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 2 -o gpurun_out/${TAG}_prof -f \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/${TAG}_ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
