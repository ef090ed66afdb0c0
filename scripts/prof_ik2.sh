#!/bin/bash
# ncu --set full capture (with source) of the fused IK kernel: one launch of the direct mode, 16384 frames
mkdir -p gpurun_out
cat > /tmp/prof_ik2.py <<'PY'
import os, sys, torch
sys.path.insert(0, ".")
import bench_ik
from smplpp_b200 import api, synth
dev = torch.device("cuda", 0)
params = synth.make_smpl_params(0)
smpl = api.SMPL(params, device=dev)
vposer = api.VPoserDecoder(synth.make_vposer_params(1), device=dev)
_, face_idx, _ = synth.make_marker_tasks(params)
tasks = api.IkTaskSet(smpl, face_idx, vposer=vposer)
prob = bench_ik.make_problem(smpl, tasks, 16384, 20, dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "direct"
opt = api.ik_options(enable_vposer=1 if mode == "vposer" else 0)
theta = prob["x0"].clone() if mode != "vposer" else torch.zeros((16384, 44), device=dev)
if mode == "vposer":
    theta[:, :6] = prob["x0"][:, :6]
vw = prob["w0"].clone()
for _ in range(4):
    tasks.step(opt, theta, prob["beta"], vw, prob["target"], pos_task_weight=prob["valid"])
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:"ik_fused" -s 2 -c 1 -o gpurun_out/prof_ik2 -f python /tmp/prof_ik2.py ${1:-direct} > gpurun_out/ncu_ik2.log 2>&1
tail -3 gpurun_out/ncu_ik2.log
