#!/bin/bash
echo "== TMA sum: parity"; AG_SUM_TMA=1 timeout 600 python -m pytest tests/test_gpu_sum.py tests/test_golden_vectors.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -3
for v in 0 1 0 1; do echo "== AG_SUM_TMA=$v"; AG_SUM_TMA=$v timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); o=d['others']; print('sum_f64 us', round(o['sum_f64']['ms']*1e3,2), round(o['sum_f64']['frac'],4), 'sum_i64 us', round(o['sum_i64']['ms']*1e3,2), round(o['sum_i64']['frac'],4))"; done
