#!/bin/bash
# Round-2 GPU call: GPU tests, both bench arms as the driver runs them, optional ncu launch list / full captures.
#   scripts/gpu_round2.sh [noprof|launches|full] [pytest -k expression]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/gpu.txt
if [ -n "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -k "$2" 2>&1 | tail -25 | tee gpurun_out/pytest.log
else
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest.log
fi
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 600 gpurun_out/bench_ref.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" == "launches" ] || [ "$1" == "full" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --passes 2 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
fi
if [ "$1" == "full" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_skin_tc3" -s 4 -c 1 -o gpurun_out/prof_fwd -f python bench.py --steps 1 --warmup 3 --passes 2 --no-cpu-baseline --no-ik > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lbs_tc_kernel" -s 2 -c 1 -o gpurun_out/prof_lbs_tc -f python bench.py --steps 1 --warmup 3 --passes 2 --no-cpu-baseline --no-ik >> gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ik_" -s 6 -c 3 -o gpurun_out/prof_ik -f python bench.py --steps 1 --warmup 3 --passes 2 --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
fi
ls -la gpurun_out | head -40
