#!/bin/bash
# what the driver runs at round end on one GPU: the GPU tests, smoke(), the default bench line
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err || tail -20 gpurun_out/final_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/final_bench.json'))
print(d['metric'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])
PY
