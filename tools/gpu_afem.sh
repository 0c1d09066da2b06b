#!/bin/bash
python -m pytest tests/test_gpu_afem.py tests/test_gpu_dist.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -30
