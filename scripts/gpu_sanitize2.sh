#!/bin/bash
# compute-sanitizer over the kernels added after the core path: casts, min/max, cumulative sum (shared-memory
# transposes + look-back => racecheck), device hand-off.  Small sizes only.
TAG=${1:-san2}
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_cast.py tests/test_minmax.py tests/test_cumsum.py tests/test_cdata.py -m gpu -q --no-header -p no:cacheprovider -x \
  -k "known_answers or safe_bounds or (min_max and (uint8 or int64 or errors)) or reference_vectors or chunked_state or cdata or roundtrip or pinned or rejected or host_api" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/${TAG}_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_cumsum.py tests/test_minmax.py -m gpu -q --no-header -p no:cacheprovider -x \
  -k "reference_vectors or chunked_state or (min_max and int32)" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -6 gpurun_out/${TAG}_racecheck.log
