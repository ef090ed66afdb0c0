#!/bin/bash
# ncu full capture of the standalone skinning kernels (TMA pipeline, then the register kernel)
mkdir -p gpurun_out
python scripts/lbs_debug.py 4096 20
ncu --set full --clock-control none --import-source on -k regex:"lbs_" -s 2 -c 1 -o gpurun_out/prof_lbs_tma python scripts/lbs_debug.py 4096 2 > gpurun_out/ncu_lbs.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lbs_kernel" -s 2 -c 1 -o gpurun_out/prof_lbs_reg python scripts/lbs_debug.py 4096 2 >> gpurun_out/ncu_lbs.log 2>&1
tail -2 gpurun_out/ncu_lbs.log | cut -c1-200
