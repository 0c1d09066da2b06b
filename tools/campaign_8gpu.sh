#!/bin/bash
# Round-2 measurements on 8 GPUs of one box
run() { # n op extra...
  n=$1; op=$2; shift; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $n --steps 30 --warmup 3 --op $op "$@" > gpurun_out/r02_scale_${op}_n$n.json 2> gpurun_out/r02_scale_${op}_n$n.err || tail -5 gpurun_out/r02_scale_${op}_n$n.err
}
run 8 stiffness
run 4 stiffness --no-e2e
run 2 stiffness --no-e2e
run 8 mass --no-e2e
run 8 general
run 8 weighted
run 8 load
run 8 eta
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29911 tools/sweep.py > gpurun_out/r02_sweep_8gpu.json 2> gpurun_out/r02_sweep_8gpu.err || tail -5 gpurun_out/r02_sweep_8gpu.err
python - <<PY
import json,glob
for p in sorted(glob.glob("gpurun_out/r02_scale_*.json")):
    try:
        t=open(p).read().split("\n"); d=json.loads([l for l in t if l.startswith("{")][0])
        print(p.split("/")[-1], round(d["ms_per_step"],3), round(d["value"]), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()}, d["e2e"] and (round(d["e2e"]["ms_per_step"],1), round(d["e2e"]["pinned"]["ms_per_step"],1)))
    except Exception as ex: print(p, "ERR", ex)
PY
grep level gpurun_out/r02_sweep_8gpu.err | cut -c1-220
