#!/bin/bash
# ncu evidence for the default path (one GPU): launch list + full capture of the hot kernels
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:scatter_kernel|fill_rows_kernel|rowptr_kernel' -s 8 -c 4 \
    -o gpurun_out/r02_default python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu1.log gpurun_out/ncu2.log
