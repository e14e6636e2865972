#!/bin/bash
# GPU round: parity suite + bench (both arms) + ncu full captures of every hot kernel.
TAG=${1:-r1c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python scripts/prof_kernels.py > gpurun_out/${TAG}_kernels.txt 2>&1; cat gpurun_out/${TAG}_kernels.txt
for thr in 0 64 2000; do echo "dense_threshold=$thr"; AG_FILTER_DENSE_THRESHOLD=$thr timeout 300 python scripts/prof_kernels.py 2>&1 | grep filter; done | tee gpurun_out/${TAG}_filter_thr.txt
timeout 900 python bench.py --steps 50 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 5 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sum_kernel|binary_vec|compare_kernel|filter_kernel|take_kernel' -c 14 -o gpurun_out/${TAG}_kernels -f \
    python scripts/prof_kernels.py 100000000 1 > gpurun_out/${TAG}_ncu_kernels.log 2>&1
echo "ncu kernels rc=$?"
