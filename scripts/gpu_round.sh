#!/bin/bash
# One GPU round: tests, bench, ncu launch list, ncu full capture of the dominant kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" != "noprof" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
fi
if [ "$1" == "full" ]; then
ncu --set full --clock-control none --import-source on -k regex:"blend_skin_tc3" -s 4 -c 2 -o gpurun_out/prof_fwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ik > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lbs_tc_kernel" -s 2 -c 1 -o gpurun_out/prof_lbs_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ik >> gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ik_jacobian|ik_solve" -s 2 -c 2 -o gpurun_out/prof_ik python bench.py --steps 2 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"vposer_jac_tc|vposer_decode|closest_point" -s 3 -c 3 -o gpurun_out/prof_aux python bench.py --steps 2 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
fi
ls -la gpurun_out
