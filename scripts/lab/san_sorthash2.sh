#!/bin/bash
TAG=${1:-san4}
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python scripts/lab/san_sorthash.py 9001 > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/${TAG}_racecheck.log
timeout 150 python -m pytest tests/test_set_lookup.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${TAG}_pytest_hash.txt 2>&1; echo "hash rc=$?"; tail -3 gpurun_out/${TAG}_pytest_hash.txt
timeout 100 python scripts/lab/r2b_lab.py 100000000 5 2>&1 | grep -E "is_in|unique" | tee gpurun_out/${TAG}_lab.txt
