#!/bin/bash
# one GPU call: the cold-path tests, then the stiffness bench on S(9)..S(11) (graph route) and S(12)
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cold" 2>&1 | tail -8
for lv in 9 10 11; do
  timeout 300 python bench.py --level $lv --steps 50 --warmup 5 --no-e2e --no-cpu > gpurun_out/g$lv.json 2> gpurun_out/g$lv.err || tail -5 gpurun_out/g$lv.err
done
timeout 300 python bench.py --level 9 --steps 50 --warmup 5 --no-e2e --no-cpu --op mass > gpurun_out/g9_mass.json 2> gpurun_out/g9_mass.err || tail -5 gpurun_out/g9_mass.err
timeout 600 python bench.py --steps 20 --no-cpu > gpurun_out/g12.json 2> gpurun_out/g12.err || tail -20 gpurun_out/g12.err
python - <<PY
import json
for n in ["g9","g10","g11","g9_mass","g12"]:
    try:
        d=json.load(open("gpurun_out/%s.json"%n)); print(n, round(d["ms_per_step"],4), round(d["value"]), d["gpu_launches"], {k:round(v["ms_per_step"],4) for k,v in d["roofline"]["kernels"].items()})
    except Exception as ex: print(n, "ERR", ex)
PY
