#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "general or weighted or cold or fullsize or window" 2>&1 | tail -3
for op in general weighted; do
python bench.py --steps 20 --op $op --no-e2e --no-cpu 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$op', round(d['ms_per_step'],4), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
done
