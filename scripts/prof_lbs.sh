#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"lbs_kernel" -s 4 -c 1 -o gpurun_out/prof_lbs python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ik > gpurun_out/ncu_lbs.log 2>&1
tail -2 gpurun_out/ncu_lbs.log | cut -c1-200
