#!/bin/bash
# launch list of the default bench step (one ncu pass, after the plain command exited 0) + the general line
python bench.py --steps 20 --op general > gpurun_out/r02_bench_general_64M.json 2> gpurun_out/r02_bench_general.err || tail -5 gpurun_out/r02_bench_general.err
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
