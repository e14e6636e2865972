#!/bin/bash
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cumsum.py tests/test_set_lookup.py tests/test_gpu_compute_api.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.txt
timeout 300 python scripts/lab/r2b_lab.py 100000000 5 2>&1 | grep -E "cumsum|is_in|unique" | tee gpurun_out/${TAG}_lab.txt
