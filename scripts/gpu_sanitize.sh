#!/bin/bash
# compute-sanitizer over the small-size parity tests (memcheck finds out-of-bounds edge-word accesses,
# racecheck shared-memory hazards in the filter kernels).  Large-row tests are deselected.
TAG=${1:-san}
mkdir -p gpurun_out
SEL='not 100m and not 125m and not properties and not reference_kernel_sweep and not random_all_types and not matches_reference_simd'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_filter.py tests/test_gpu_bitmap.py tests/test_gpu_take.py tests/test_gpu_compute_api.py tests/test_golden_vectors.py -m gpu -q --no-header -p no:cacheprovider -x -k "$SEL" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -8 gpurun_out/${TAG}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_filter.py -m gpu -q --no-header -p no:cacheprovider -x -k "literal or fused or take_indices" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -8 gpurun_out/${TAG}_racecheck.log
