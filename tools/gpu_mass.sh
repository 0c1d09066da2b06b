#!/bin/bash
# one GPU call: parity tests, then the mass matrix (16-byte slots) and the stiffness at 64M
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -4
for op in mass stiffness; do
python bench.py --steps 20 --op $op --no-e2e --no-cpu 2>gpurun_out/m_$op.err | grep '^{' > gpurun_out/m_$op.json || tail -5 gpurun_out/m_$op.err
python -c "
import json
d=json.load(open('gpurun_out/m_$op.json')); print('$op', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
done
python bench.py --steps 50 --op mass --level 9 --no-e2e --no-cpu 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('mass S9', round(d['ms_per_step'],4))"
