#!/bin/bash
# one 8-GPU box, final code: stiffness at N = 1, 2, 4, 8 back to back (what the driver's scaling run does), then the sweep at N = 8
python bench.py --steps 20 --no-e2e --no-cpu 2>/dev/null | grep '^{' > gpurun_out/r02_final_stiffness_n1.json
for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 30 --warmup 3 --no-e2e 2>/dev/null | grep '^{' > gpurun_out/r02_final_stiffness_n$n.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29790 tools/sweep.py 9 10 11 12 13 > gpurun_out/r02_sweep_8gpu.json 2> gpurun_out/r02_sweep_8gpu.err || tail -5 gpurun_out/r02_sweep_8gpu.err
python - <<PY
import json
base=None
for n in (1,2,4,8):
    d=json.load(open("gpurun_out/r02_final_stiffness_n%d.json"%n))
    base=base or d["value"]
    print(n, round(d["ms_per_step"],4), round(d["value"]), round(d["value"]/base,2), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()})
t=open("gpurun_out/r02_sweep_8gpu.json").read(); d=json.loads(t[t.index("{"):])
for r in d["rows"]: print(r["level"], round(r["stiffness_to_csr_ms"],4), round(r["stiffness_Gelem_s"],1), round(r["eta_r_cold_ms"],4))
PY
