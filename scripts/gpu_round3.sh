#!/bin/bash
# Final GPU call of round 2: GPU tests, both bench arms as the driver runs them, ncu launch lists (bench command and the
# IK iteration) and --set full captures of the dominant kernels (forward, standalone skinning, the four IK kernels).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 300 gpurun_out/bench_ref.json; echo
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); r=d['roofline']
print('value %.3f M meshes/s, ms/launch %.4f (burst %.4f), frac %.3f, e2e %.0f, ik %.3f M (%.3f ms), vposer %.3f M, shared-beta %.3f M, cfg3 %.3f M, jacobian %.2f ms' % (d['value']/1e6, r['ms_per_launch'], r['burst_ms_per_launch'], r['frac'], d['e2e']['value'], d['ik_value']/1e6, d['ik']['mosh_direct']['ms_per_iter'], d['ik_moshpp_vposer']/1e6, d['ik_shared_beta']/1e6, d['ik_shared_beta_vposer']/1e6, d['ik']['jacobian']['ms_per_call']))"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --passes 2 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
bash scripts/launch_list_ik.sh > gpurun_out/launch_list_ik.txt 2>&1; tail -8 gpurun_out/launch_list_ik.txt
cat > /tmp/prof_fwd.py <<'PY'
import sys, ctypes as C, torch
sys.path.insert(0, ".")
from smplpp_b200 import api, capi, synth
dev = torch.device("cuda", 0)
params = synth.make_smpl_params(0)
smpl = api.SMPL(params, device=dev)
B = 4096
beta_h, theta_h = synth.make_forward_inputs(B, 11)
for _ in range(6):
    smpl.launch(beta_h, theta_h)
torch.cuda.synchronize()
lib = capi.lib()
beta, theta = torch.as_tensor(beta_h, device=dev), torch.as_tensor(theta_h, device=dev)
ws_bytes = lib.smplpp_forward_workspace_bytes(smpl.handle, C.c_int64(B))
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
rest = torch.empty((B, 6890, 3), device=dev); xf = torch.empty((B, 24, 4, 4), device=dev); verts = torch.empty((B, 6890, 3), device=dev)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
capi.check(lib.smplpp_forward(smpl.handle, st, C.c_int64(B), C.c_void_p(beta.data_ptr()), C.c_int64(10), C.c_void_p(theta.data_ptr()), None, None,
                              C.c_void_p(xf.data_ptr()), C.c_void_p(rest.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_size_t(ws_bytes)))
xf34 = xf[:, :, :3, :].contiguous(); root = theta[:, 0].contiguous()
for _ in range(4):
    capi.check(lib.smplpp_model_skinning34(smpl.handle, st, C.c_int64(B), C.c_void_p(rest.data_ptr()), C.c_void_p(xf34.data_ptr()),
                                           C.c_void_p(root.data_ptr()), C.c_void_p(verts.data_ptr())))
torch.cuda.synchronize()
PY
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"blend_skin_tc3" -s 3 -c 1 -o gpurun_out/prof_fwd -f python /tmp/prof_fwd.py > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lbs_tc_kernel" -s 2 -c 1 -o gpurun_out/prof_lbs_tc -f python /tmp/prof_fwd.py >> gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ik_" -s 4 -c 4 -o gpurun_out/prof_ik -f python /tmp/prof_ik.py >> gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep | tail -5
