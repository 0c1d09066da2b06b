#!/bin/bash
# one GPU call: the GPU test suite, every operator at S(9), the default bench at S(12)
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
for op in stiffness mass general weighted load eta; do
  python bench.py --level 9 --steps 5 --op $op > gpurun_out/b9_$op.json 2> gpurun_out/b9_$op.err || tail -5 gpurun_out/b9_$op.err
done
python bench.py --steps 20 > gpurun_out/b12.json 2> gpurun_out/b12.err || tail -20 gpurun_out/b12.err
python - <<PY
import json
for op in ["stiffness","mass","general","weighted","load","eta"]:
    try:
        d=json.load(open("gpurun_out/b9_%s.json"%op)); print(op, round(d["ms_per_step"],4), round(d["value"]), d["e2e"] and round(d["e2e"]["ms_per_step"],2), d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2), d["clocks"]["samples"])
    except Exception as ex: print(op, "ERR", ex)
d=json.load(open("gpurun_out/b12.json")); print(d["ms_per_step"], d["roofline"]["frac"], {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()}, d["e2e"], d["cpu_baseline"], d["clocks"])
PY
