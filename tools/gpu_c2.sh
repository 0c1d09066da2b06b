#!/bin/bash
# config C2: S(11) = 16M triangles, mass matrix and cubature load vector, one GPU
for op in mass load; do
python bench.py --level 11 --steps 20 --op $op > gpurun_out/r02_bench_${op}_16M.json 2> gpurun_out/r02_bench_${op}_16M.err || tail -5 gpurun_out/r02_bench_${op}_16M.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_${op}_16M.json')); print('$op', round(d['ms_per_step'],4), round(d['value']), round(d['roofline']['frac'],4), round(d['e2e']['ms_per_step'],2), round(d['e2e']['pinned']['ms_per_step'],2), round(d['cpu_baseline']['value'],2))"
done
