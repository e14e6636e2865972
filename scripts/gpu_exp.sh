#!/bin/bash
for v in 1 0; do
  echo "== AG_EW_VEC=$v"
  AG_EW_VEC=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-others --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunked ms', round(d['ms_per_step'],4), 'contig', round(d['config']['contiguous_ms_per_step'],4))"
done
