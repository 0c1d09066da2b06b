#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29771 tools/exchange_probe.py > gpurun_out/r02_exchange_probe.json 2> gpurun_out/r02_exchange_probe.err || tail -5 gpurun_out/r02_exchange_probe.err
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 50 --warmup 3 --no-e2e > gpurun_out/r02_final_stiffness_n$n.json 2>/dev/null
done
python - <<PY
import json
print(open("gpurun_out/r02_exchange_probe.json").read()[-700:])
for n in (2,4,8):
    t=open("gpurun_out/r02_final_stiffness_n%d.json"%n).read().split("\n"); d=json.loads([l for l in t if l.startswith("{")][0])
    print(n, round(d["ms_per_step"],4), round(d["value"]), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()})
PY
