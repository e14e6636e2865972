#!/bin/bash
# cast round: parity of the numeric-cast kernels + promotion through the host API + kernel timings
TAG=${1:-c1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cast.py tests/test_gpu_compute_api.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt
timeout 300 python scripts/prof_kernels.py 2>&1 | tee gpurun_out/${TAG}_kernels.txt
