#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 100 python scripts/lab/r2b_lab.py 100000000 5 2>&1 | grep -E "is_in|unique|cumsum" | tee gpurun_out/final_lab.txt
