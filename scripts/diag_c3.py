import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from smplpp_b200 import api, synth
import test_ik_configs_gpu as t
gc = dict(np.load("tests/golden/ref_ik_configs.npz"))
params = synth.make_smpl_params(0)
smpl = api.SMPL(params, device="cuda:0")
_, face_idx, _ = synth.make_marker_tasks(params)
ts = api.IkTaskSet(smpl, face_idx, vposer=api.VPoserDecoder(synth.make_vposer_params(1)))
np.set_printoptions(linewidth=250, precision=2, suppress=False)
opt = api.ik_options(**t.MOTION)
res, traj, vw = t.run_trajectory(ts, opt, gc["c3_theta_in"], gc["beta"], gc["vertex_weights_in"], gc["c3_target"], gc["c3_valid"], 30)
d = np.abs(res - gc["c3_residual"])
print("c3 residual diff per frame (max over iters):", d.max(1))
print("argmax iter per frame:", d.argmax(1))
dth = np.abs(traj - gc["c3_theta_traj"]).max(2)
print("theta diff per frame/iter:\n", dth)
print("golden residual:\n", gc["c3_residual"])
optv = api.ik_options(enable_vposer=1, **t.MOTION)
res, traj, _ = t.run_trajectory(ts, optv, gc["c4_theta_in"], gc["beta"], gc["vertex_weights_in"], gc["c4_target"], gc["c4_valid"], 10)
print("c4 residual diff:\n", np.abs(res - gc["c4_residual"]))
print("c4 theta diff:\n", np.abs(traj - gc["c4_theta_traj"]).max(2))
