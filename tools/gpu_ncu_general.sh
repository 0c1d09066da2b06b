#!/bin/bash
# ncu --set full of the scatter of the generalised stiffness and of the weighted mass (after the plain commands exited 0)
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --op general > gpurun_out/plain_g.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:scatter_kernel_wide' -s 1 -c 1 \
    -o gpurun_out/r02_general python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --op general > gpurun_out/ncu_g.log 2>&1
tail -2 gpurun_out/ncu_g.log; ls -la gpurun_out/r02_general.ncu-rep
