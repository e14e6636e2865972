#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_parquet.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/r2x_pytest_parquet.txt 2>&1; echo "parquet rc=$?"; tail -2 gpurun_out/r2x_pytest_parquet.txt
timeout 60 python scripts/lab/parquet_time.py 2>&1 | tee gpurun_out/r2x_parquet_time.txt
timeout 300 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2x_pytest_all.txt 2>&1; echo "all rc=$?"; tail -2 gpurun_out/r2x_pytest_all.txt
