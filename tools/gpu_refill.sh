#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_afem.py -x -q -m gpu 2>&1 | tail -3
python bench_ops.py --level 12 --cpu-level 8 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d['ops'].items(): print(round(v['gpu_ms'],3), k[:110])"
