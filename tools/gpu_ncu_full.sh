#!/bin/bash
# ncu --set full of the kernels of one cold step (after the plain command exited 0)
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:scatter_kernel|fill_rows_kernel|rowptr_kernel|cold_init' -s 12 -c 6 \
    -o gpurun_out/r02_default python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log; ls -la gpurun_out/r02_default.ncu-rep
