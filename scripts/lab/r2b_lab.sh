#!/bin/bash
# usage: bash scripts/lab/r2b_lab.sh <tag>
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cumsum.py tests/test_gpu_filter.py tests/test_gpu_take.py tests/test_gpu_compute_api.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt
timeout 600 python scripts/lab/r2b_lab.py 2>&1 | tee gpurun_out/${TAG}_lab.txt
