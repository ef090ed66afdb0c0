#!/bin/bash
mkdir -p gpurun_out
for aff in 1 0; do
  if [ $aff == 0 ]; then export SMPLPP_BENCH_NO_AFFINITY=1; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$aff bench.py --gpus 2 --steps 20 --warmup 5 --no-ik 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('affinity', $aff, 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('cpu_affinity_cores'), 'pageable', d['e2e']['pageable_value'])"
done
nvidia-smi topo -m 2>/dev/null | head -12
