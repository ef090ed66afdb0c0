#!/bin/bash
# 2-GPU validation of round 2: the C entry with a raw ncclComm_t, gather over NCCL, the driver's launch line with the
# configs[4] leg (2^20 frames sharded over the ranks).
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/nccl_c_entry_check.py 2>&1 | grep -E "rank [01]/|Error|error" | tee gpurun_out/nccl_c_entry_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n2.json"))
print(json.dumps({k: d[k] for k in d if k.startswith("ik_") or k in ("value", "n_gpus")}))
print(json.dumps(d["ik"]["config4"]))
print(json.dumps(d["e2e"]))
PY
