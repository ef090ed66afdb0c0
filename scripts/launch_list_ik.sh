# per-kernel device times of the IK step (direct mode, 16384 frames): ncu launch list
cat > /tmp/ll_ik.py <<'PY'
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
for _ in range(4):
    tasks.step(opt, theta, prob["beta"], vw, prob["target"], pos_task_weight=prob["valid"])
torch.cuda.synchronize()
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ik.csv python /tmp/ll_ik.py > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.DictReader(l for l in open("gpurun_out/launches_ik.csv") if l.startswith('"'))]
rows = [r for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"]
for r in rows[-12:]:
    print("%-60s %10s %s" % (r["Kernel Name"][:60], r["Metric Value"], r["Metric Unit"]))
PY
