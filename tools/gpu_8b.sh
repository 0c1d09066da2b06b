#!/bin/bash
# 8 GPUs: the distributed tests (2-rank NCCL included), then load / stiffness / mass at N = 8
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -3
for op in load stiffness; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --op $op --level 12 --no-e2e > gpurun_out/n8_$op.out 2> gpurun_out/n8_$op.err || tail -20 gpurun_out/n8_$op.err
grep '^{' gpurun_out/n8_$op.out > gpurun_out/n8_$op.json
python -c "
import json
d=json.load(open('gpurun_out/n8_$op.json')); print('$op', d['ms_per_step'], round(d['value']), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
done
