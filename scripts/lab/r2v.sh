#!/bin/bash
TAG=${1:-r2v}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_cumsum.py tests/test_gpu_compute_api.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${TAG}_pytest.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/${TAG}_pytest.txt
timeout 100 python scripts/lab/r2b_lab.py 100000000 5 2>&1 | grep -E "cumsum" | tee gpurun_out/${TAG}_lab.txt
