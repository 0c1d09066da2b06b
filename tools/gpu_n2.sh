#!/bin/bash
# 2 GPUs: distributed tests, then every operator of bench.py at N = 2 (device-resident legs only)
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -3
for op in stiffness mass general weighted load eta; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --op $op --level 12 --no-e2e > gpurun_out/n2_$op.out 2> gpurun_out/n2_$op.err || tail -20 gpurun_out/n2_$op.err
grep '^{' gpurun_out/n2_$op.out > gpurun_out/n2_$op.json
python -c "
import json
d=json.load(open('gpurun_out/n2_$op.json')); print('$op', round(d['ms_per_step'],4), round(d['value']), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
done
