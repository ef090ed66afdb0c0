#!/bin/bash
# ncu evidence of round 2: launch list of a short bench run + --set full captures of the dominant kernels.
mkdir -p gpurun_out
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
# standalone skinning (lbs_tc_kernel): rest shape + transforms -> vertices
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
vposer = api.VPoserDecoder(synth.make_vposer_params(1), device=dev)
_, face_idx, _ = synth.make_marker_tasks(params)
tasks = api.IkTaskSet(smpl, face_idx, vposer=vposer)
prob = bench_ik.make_problem(smpl, tasks, 16384, 20, dev)
for variant in (401, 402):
    capi.check(capi.lib().smplpp_set_forward_variant(variant))
    opt = api.ik_options()
    theta, vw = prob["x0"].clone(), prob["w0"].clone()
    for _ in range(3):
        tasks.step(opt, theta, prob["beta"], vw, prob["target"], pos_task_weight=prob["valid"])
torch.cuda.synchronize()
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --passes 4 --no-cpu-baseline --no-config4 > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"blend_skin_tc3" -s 3 -c 1 -o gpurun_out/prof_fwd -f python /tmp/prof_fwd.py > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lbs_tc_kernel" -s 2 -c 1 -o gpurun_out/prof_lbs_tc -f python /tmp/prof_fwd.py >> gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ik_jacobian|ik_solve|ik_fused" -s 4 -c 4 -o gpurun_out/prof_ik -f python /tmp/prof_ik.py >> gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
