#!/bin/bash
# N=4 sanity run of both bench arms with the driver's launch line
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; tail -3 gpurun_out/bench_n4.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_ref_n4.json 2> gpurun_out/bench_ref_n4.err; tail -2 gpurun_out/bench_ref_n4.err
wc -l gpurun_out/bench_n4.json gpurun_out/bench_ref_n4.json
cut -c1-600 gpurun_out/bench_n4.json
