#!/bin/bash
# ncu full capture of the tcgen05 blend + tensor-core skinning kernel + plain timing of all variants
mkdir -p gpurun_out
timeout 300 python scripts/tc_debug.py 4096 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"blend_skin_tc2" -s 5 -c 1 -o gpurun_out/prof_tc2 python scripts/tc_debug.py 4096 > gpurun_out/ncu_tc.log 2>&1
tail -3 gpurun_out/ncu_tc.log | cut -c1-200
