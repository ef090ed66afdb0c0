#!/bin/bash
# N=8 sanity run of the bench with the driver's launch line
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -2 gpurun_out/bench_n8.err
wc -l gpurun_out/bench_n8.json; cut -c1-300 gpurun_out/bench_n8.json
