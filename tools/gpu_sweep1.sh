#!/bin/bash
# one GPU call: the C5 sweep on one GPU, the S(9) bench line, the launch list of the S(9) step
python tools/sweep.py 9 10 11 12 13 > gpurun_out/sweep_1gpu.json 2> gpurun_out/sweep_1gpu.err || tail -5 gpurun_out/sweep_1gpu.err
python bench.py --level 9 --steps 100 --warmup 10 > gpurun_out/b9.json 2> gpurun_out/b9.err || tail -5 gpurun_out/b9.err
python bench.py --steps 20 > gpurun_out/b12.json 2> gpurun_out/b12.err || tail -5 gpurun_out/b12.err
python - <<PY
import json
d=json.load(open('gpurun_out/sweep_1gpu.json'))
for r in d['rows']: print(r['level'], round(r['stiffness_to_csr_ms'],4), round(r['stiffness_to_csr_same_arrays_ms'],4), round(r['eta_r_cold_ms'],4))
for n in ('b9','b12'):
    d=json.load(open('gpurun_out/%s.json'%n)); print(n, d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['ms_per_step'], d['e2e']['pinned']['ms_per_step'])
PY
