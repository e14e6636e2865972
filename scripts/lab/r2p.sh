#!/bin/bash
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_sort.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${TAG}_pytest_sort.txt 2>&1; echo "sort rc=$?"; tail -3 gpurun_out/${TAG}_pytest_sort.txt
timeout 100 python scripts/sort_bench.py 100000000 2>&1 | tee gpurun_out/${TAG}_sort.txt
