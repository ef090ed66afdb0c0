# ncu source-level capture of the IK kernels named by the regex in argument 1 (direct mode, 16384 frames)
cat > /tmp/prof_ik.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bench_ik
from smplpp_b200 import api, capi, synth
dev = torch.device("cuda", 0)
params = synth.make_smpl_params(0)
smpl = api.SMPL(params, device=dev)
_, face_idx, _ = synth.make_marker_tasks(params)
tasks = api.IkTaskSet(smpl, face_idx)
prob = bench_ik.make_problem(smpl, tasks, 16384, 20, dev)
opt = api.ik_options()
theta, vw = prob["x0"].clone(), prob["w0"].clone()
for _ in range(3):
    tasks.step(opt, theta, prob["beta"], vw, prob["target"], pos_task_weight=prob["valid"])
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$1" -s 2 -c 1 -o gpurun_out/prof_ik_src -f python /tmp/prof_ik.py > gpurun_out/ncu_ik_src.log 2>&1
tail -2 gpurun_out/ncu_ik_src.log
