#!/bin/bash
TAG=${1:-san3}
mkdir -p gpurun_out
timeout 240 python scripts/lab/san_sorthash.py > gpurun_out/${TAG}_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/${TAG}_plain.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python scripts/lab/san_sorthash.py 20011 > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/${TAG}_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python scripts/lab/san_sorthash.py 9001 > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/${TAG}_racecheck.log
