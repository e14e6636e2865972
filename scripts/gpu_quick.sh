#!/bin/bash
# quick experiment round: filter/take parity + kernel timings under the L2 fetch granularity knob
TAG=${1:-q1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_filter.py tests/test_gpu_take.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -4
for g in default 32 64 128; do
  echo "== L2 fetch granularity: $g"
  if [ "$g" = default ]; then timeout 300 python scripts/prof_kernels.py; else AG_L2_FETCH_GRANULARITY=$g timeout 300 python scripts/prof_kernels.py; fi
done 2>&1 | tee gpurun_out/${TAG}_granularity.txt
for thr in 0 2000; do echo "dense_threshold=$thr (granularity 32)"; AG_L2_FETCH_GRANULARITY=32 AG_FILTER_DENSE_THRESHOLD=$thr timeout 300 python scripts/prof_kernels.py 2>&1 | grep filter; done | tee -a gpurun_out/${TAG}_granularity.txt
