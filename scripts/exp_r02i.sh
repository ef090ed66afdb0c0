# one call: IK tests + quick IK bench of the working tree, TMEM / MMA micro-benchmarks
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "ik or shared_beta or task" 2>&1 | tail -4
timeout 300 python scripts/bench_ik_quick.py 2>&1 | tail -3
timeout 120 scripts/ubench/ubench 2>&1 | tee gpurun_out/ubench_tmem.log | head -40
timeout 120 scripts/ubench/ubench_mma 2>&1 | tee gpurun_out/ubench_mma.log | tail -60
