#!/bin/bash
TAG=${1:-s1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cumsum.py tests/test_minmax.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest.txt
timeout 300 python scripts/prof_kernels.py 2>&1 | grep -E "cumsum|min_max|sum_f64" | tee gpurun_out/${TAG}_kernels.txt
