#!/bin/bash
# take: parity + timings + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_take.py tests/test_gpu_arith.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python scripts/take_sweep.py 100000000,100000000 1000000000,125000000 100000000,25000000 2>&1 | tail -4
timeout 300 python scripts/misaligned_bench.py 2>&1 | tail -25
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:take_ --csv --log-file gpurun_out/take_launches.csv python scripts/take_once.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/take_launches.csv')) if len(r)>10]
h=rows[0]; acc={}
for r in rows[1:]:
    d=dict(zip(h,r)); acc.setdefault((d['ID'],d['Kernel Name'][:50]),{})[d['Metric Name']]=d['Metric Value']
for k,v in list(acc.items())[-6:]: print(k[1], v)
PY
