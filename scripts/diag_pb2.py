import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from smplpp_b200 import api, capi, synth
import test_ik_configs_gpu as T
np.set_printoptions(linewidth=200)
gc = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_ik_configs.npz")))
params = synth.make_smpl_params(0)
smpl = api.SMPL(params, device="cuda:0")
_, face_idx, _ = synth.make_marker_tasks(params)
ts = api.IkTaskSet(smpl, face_idx, vposer=api.VPoserDecoder(synth.make_vposer_params(1)))
opt = api.ik_options(enable_vposer=1, **T.MOTION)
K = gc["c4_residual"].shape[1]
print("ref", np.array2string(gc["c4_residual"], precision=5))
for v in (421, 420):
    capi.check(capi.lib().smplpp_set_forward_variant(v))
    res, traj, _ = T.run_trajectory(ts, opt, gc["c4_theta_in"], gc["beta"], gc["vertex_weights_in"], gc["c4_target"], gc["c4_valid"], K)
    dev = np.abs(res - gc["c4_residual"])
    print(v, np.array2string(res, precision=5))
    print(v, "dev[:, :4].max %.3g rel max %.3g theta0 %.3g" % (dev[:, :4].max(), (dev / gc["c4_residual"]).max(), np.abs(traj[:, 0] - gc["c4_theta_traj"][:, 0]).max()))
# teacher-forced J comparison in VPoser mode
th, vw, res = gc["c4_theta_traj"], gc["c4_vertex_weights_traj"], gc["c4_residual"]
F, K = res.shape; n = vw.shape[2]
th_in = np.concatenate([np.repeat(gc["c4_theta_in"][None, None], F, axis=0), th[:, :-1]], axis=1)
vw_in = np.concatenate([np.repeat(gc["vertex_weights_in"][None, None], F, axis=0), vw[:, :-1]], axis=1)
tgt = T.cu(np.repeat(gc["c4_target"][:, None], K, axis=1).reshape(F * K, n, 3))
valid = np.repeat(gc["c4_valid"][:, None], K, axis=1).reshape(F * K, n)
Js = {}
for v in (421, 420):
    capi.check(capi.lib().smplpp_set_forward_variant(v))
    theta, w = T.cu(th_in.reshape(F * K, 44)), T.cu(vw_in.reshape(F * K, n, 3))
    status, out = ts.step(opt, theta, T.cu(gc["beta"]), w, tgt, pos_task_weight=T.cu(valid), outputs=True)
    Js[v] = out["J"].cpu().numpy().astype(np.float64)
d = np.abs(Js[420] - Js[421])
print("VPoser J tc vs ffma: max abs %.3g, max|J| %.3g" % (d.max(), np.abs(Js[421]).max()), "argmax", np.unravel_index(d.argmax(), d.shape))
