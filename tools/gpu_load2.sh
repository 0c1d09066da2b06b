#!/bin/bash
# 2 GPUs: the distributed tests, then the load vector by element blocks + all-reduce
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -4
for op in load stiffness; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --op $op --level 12 --no-e2e > gpurun_out/l2_$op.out 2> gpurun_out/l2_$op.err || tail -20 gpurun_out/l2_$op.err
grep '^{' gpurun_out/l2_$op.out | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['parallelism'][:60], d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
done
