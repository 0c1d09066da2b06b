#!/bin/bash
# Round-2 measurements on one B200 (writes gpurun_out/r02_*)
set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 20 > gpurun_out/r02_bench_line_1gpu.json 2> gpurun_out/r02_bench_1gpu.err || tail -5 gpurun_out/r02_bench_1gpu.err
for op in mass general weighted load eta; do
  python bench.py --steps 20 --op $op > gpurun_out/r02_bench_${op}_64M.json 2> gpurun_out/r02_bench_${op}.err || tail -5 gpurun_out/r02_bench_${op}.err
done
python bench.py --steps 100 --warmup 10 --level 9 > gpurun_out/r02_bench_line_1M.json 2>/dev/null
python bench_ops.py --level 12 --cpu-level 10 > gpurun_out/r02_ops_64M.json 2> gpurun_out/r02_ops.err || tail -5 gpurun_out/r02_ops.err
python tools/sweep.py > gpurun_out/r02_sweep_1gpu.json 2> gpurun_out/r02_sweep_1gpu.err || tail -5 gpurun_out/r02_sweep_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
