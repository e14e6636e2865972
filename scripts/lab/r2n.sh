#!/bin/bash
TAG=${1:-r2n}
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_set_lookup.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${TAG}_pytest_hash.txt 2>&1; echo "hash rc=$?"; tail -3 gpurun_out/${TAG}_pytest_hash.txt
timeout 200 python -m pytest tests/test_sort.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${TAG}_pytest_sort.txt 2>&1; echo "sort rc=$?"; tail -3 gpurun_out/${TAG}_pytest_sort.txt
timeout 100 python scripts/sort_bench.py 100000000 2>&1 | tee gpurun_out/${TAG}_sort.txt
timeout 100 python scripts/lab/r2b_lab.py 100000000 5 2>&1 | grep -E "is_in|unique" | tee gpurun_out/${TAG}_lab.txt
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/lab/r2k_once.py 100000000 > gpurun_out/${TAG}_once.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"hash_insert_kernel|unique_mark_kernel" -c 3 -o gpurun_out/${TAG}_prof_uniq -f python scripts/lab/r2k_once.py 32000000 > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"is_in_kernel" -c 1 -o gpurun_out/${TAG}_prof_isin -f python scripts/lab/r2k_once.py 32000000 > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
