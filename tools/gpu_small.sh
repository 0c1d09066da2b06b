#!/bin/bash
# one GPU call: cold-path tests, host/device time of a small cold assembly, its launch list
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -5
python tools/small_probe.py 9; python tools/small_probe.py 7; python tools/small_probe.py 10; python tools/small_probe.py 11
python bench.py --level 9 --steps 3 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_s9.csv python bench.py --level 9 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_s9.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/launches_s9.csv')))
h=[r for r in rows if 'Kernel Name' in r][0]; i=rows.index(h); k=h.index('Kernel Name'); v=h.index('Metric Value')
for r in rows[i+1:][-10:]:
    if len(r)>v: print(r[k].split('(')[0][:50], r[h.index('Grid Size')], r[v])
PY
