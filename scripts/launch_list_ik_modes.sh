# per-kernel device times of one IK iteration in every mode of the bench (16384 frames): ncu launch list
cat > /tmp/ll_ik2.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bench_ik
from smplpp_b200 import api, capi, synth
dev = torch.device("cuda", 0)
params = synth.make_smpl_params(0)
smpl = api.SMPL(params, device=dev)
vposer = api.VPoserDecoder(synth.make_vposer_params(1), device=dev)
_, face_idx, _ = synth.make_marker_tasks(params)
tasks = api.IkTaskSet(smpl, face_idx, vposer=vposer)
F = 16384
prob = bench_ik.make_problem(smpl, tasks, F, 20, dev)
torch.cuda.synchronize()
print("MARK direct")
opt = api.ik_options()
theta, vw = prob["x0"].clone(), prob["w0"].clone()
for _ in range(2): tasks.step(opt, theta, prob["beta"], vw, prob["target"], pos_task_weight=prob["valid"])
optv = api.ik_options(enable_vposer=1)
xv = torch.zeros((F, 44), device=dev); xv[:, :6] = prob["x0"][:, :6]
vw = prob["w0"].clone()
for _ in range(2): tasks.step(optv, xv, prob["beta"], vw, prob["target"], pos_task_weight=prob["valid"])
sb = torch.zeros(10, device=dev); th2, vw2 = prob["x0"].clone(), prob["w0"].clone()
for _ in range(2): tasks.shared_beta_step(opt, th2, sb, vw2, prob["target"], pos_task_weight=prob["valid"])
torch.cuda.synchronize()
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ik_modes.csv python /tmp/ll_ik2.py > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.DictReader(l for l in open("gpurun_out/launches_ik_modes.csv") if l.startswith('"'))]
rows = [r for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"]
# the last 2 iterations of each mode: print the tail
for r in rows[-40:]:
    print("%-70s %10s %s" % (r["Kernel Name"][:70], r["Metric Value"], r["Metric Unit"]))
PY
