#!/bin/bash
# per-phase cycle counters of the fused IK kernel (instrumented library, scripts/build_dbg.sh)
SMPLPP_B200_LIB=$PWD/smplpp_b200/libsmplpp_b200_dbg.so python - <<'PY' 2>&1 | grep "ik2 phase" | tail -4
import sys, torch
sys.path.insert(0, ".")
import bench_ik
from smplpp_b200 import api, synth, capi
capi.check(capi.lib().smplpp_set_forward_variant(402))
dev = torch.device("cuda", 0)
params = synth.make_smpl_params(0)
smpl = api.SMPL(params, device=dev)
_, face_idx, _ = synth.make_marker_tasks(params)
tasks = api.IkTaskSet(smpl, face_idx)
prob = bench_ik.make_problem(smpl, tasks, 16384, 20, dev)
opt = api.ik_options()
theta, vw = prob["x0"].clone(), prob["w0"].clone()
for _ in range(5):
    tasks.step(opt, theta, prob["beta"], vw, prob["target"], pos_task_weight=prob["valid"])
    torch.cuda.synchronize()
PY
