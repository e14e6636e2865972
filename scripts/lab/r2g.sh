#!/bin/bash
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -12 | tee gpurun_out/${TAG}_pytest.txt
timeout 300 python scripts/lab/r2b_lab.py 100000000 10 2>&1 | tee gpurun_out/${TAG}_lab.txt
