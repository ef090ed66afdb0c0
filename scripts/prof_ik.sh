#!/bin/bash
# ncu --set full capture of the two IK kernels (one launch each) from the IK leg of the bench
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"ik_jacobian|ik_solve" -s 6 -c 2 -o gpurun_out/prof_ik -f python - <<'PY' > gpurun_out/ncu_ik.log 2>&1
import sys, torch
sys.path.insert(0, '.')
from smplpp_b200 import ik_bench
dev = torch.device("cuda", 0)
ik_bench.run(dev, 0, 1, lambda x: x, torch.cuda.synchronize, iters=3, warmup=1)
PY
tail -3 gpurun_out/ncu_ik.log
