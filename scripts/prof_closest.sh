#!/bin/bash
# ncu --set full capture of the mesh-projection kernel (one launch) + its CUDA-event timing
mkdir -p gpurun_out
cat > /tmp/cp_run.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from smplpp_b200 import api, synth
params = synth.make_smpl_params(0)
smpl = api.SMPL(params, device="cuda:0")
_, face_idx, vw = synth.make_marker_tasks(params)
tasks = api.IkTaskSet(smpl, face_idx)
B = 4096
beta, theta = synth.make_forward_inputs(B, 11)
smpl.launch(beta, theta)
v = smpl.getVertex()
w = torch.as_tensor(vw, device="cuda:0")[None].repeat(B, 1, 1).contiguous()
pts = tasks.positions(v, w, 0.015)
for _ in range(3): smpl.projectPoints(pts, v)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): smpl.projectPoints(pts, v)
b.record(); torch.cuda.synchronize()
print("closest points: %.3f ms per %d frames" % (a.elapsed_time(b) / 5, B))
PY
python /tmp/cp_run.py
ncu --set full --clock-control none --import-source on -k regex:"closest_point" -s 2 -c 1 -f -o gpurun_out/prof_closest python /tmp/cp_run.py > gpurun_out/ncu_closest.log 2>&1
tail -2 gpurun_out/ncu_closest.log
