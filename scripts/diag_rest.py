"""Where do the tensor-core and FFMA variants of the two-kernel IK path differ on a random 5-face task set?"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smplpp_b200 import api, capi, synth
f32 = np.float32
cu = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda:0").contiguous()
params = synth.make_smpl_params(0)
smpl = api.SMPL(params, device="cuda:0")
n_tasks, seed = 5, 2
rng = np.random.default_rng(seed)
faces = rng.choice(params.face_indices.shape[0], size=n_tasks, replace=False).astype(np.int64)
ts = api.IkTaskSet(smpl, faces)
F = 37
theta = synth.make_motion(F, 7 + seed).reshape(F, 75).astype(f32)
theta[:, 3:] += rng.normal(scale=0.2, size=(F, 72)).astype(f32)
beta = (rng.normal(size=10) * 0.5).astype(f32)
vw = rng.dirichlet(np.ones(3), size=(F, n_tasks)).astype(f32)
tgt = rng.normal(scale=0.5, size=(F, n_tasks, 3)).astype(f32)
tn = rng.normal(size=(F, n_tasks, 3)).astype(f32); tn /= np.linalg.norm(tn, axis=2, keepdims=True)
smpl.launch(beta, theta.reshape(F, 25, 3))
V = smpl.getVertex().cpu().numpy()
fi = params.face_indices.astype(np.int64) - 1
for m, fc in enumerate(faces):
    tri = V[0][fi[fc]]
    print("task %d face %d edges (mm):" % (m, fc), np.round(1e3 * np.linalg.norm(tri - np.roll(tri, 1, 0), axis=1), 2))
for kw in (dict(normal_offset=0.015), dict(normal_task_weight=1.0, normal_offset=0.0), dict(normal_offset=0.0)):
    opt = api.ik_options(update_state=0, skip_if_too_few=0, **kw)
    outs = {}
    for variants in ((411, 421), (410, 421), (410, 420)):
        for v in variants:
            capi.check(capi.lib().smplpp_set_forward_variant(v))
        status, out = ts.step(opt, cu(theta), cu(beta), cu(vw), cu(tgt), target_normal=cu(tn) if kw.get("normal_task_weight") else None, outputs=True)
        outs[variants] = (out["e"].cpu().numpy(), out["J"].cpu().numpy())
    e0, J0 = outs[(411, 421)]
    for v in ((410, 421), (410, 420)):
        e1, J1 = outs[v]
        d = np.abs(J1 - J0)
        idx = np.unravel_index(d.argmax(), d.shape)
        print(kw, v, "max|de| %.3g  max|dJ| %.3g at (frame, row, col) %s, |J| there %.3g, max|J| %.3g" % (np.abs(e1 - e0).max(), d.max(), idx, abs(J0[idx]), np.abs(J0).max()))
print("---- same inputs: rest shape of both kernels, run-to-run determinism of e / J")
r_tc, ids = ts.restShape(beta, theta, 0)
r_ff, _ = ts.restShape(beta, theta, 1)
d = (r_tc - r_ff).abs().cpu().numpy()
print("rest tc vs ffma: max %.3g, per frame (1e-8):" % d.max(), np.round(d.reshape(F, -1).max(1) * 1e8).astype(int))
opt = api.ik_options(update_state=0, skip_if_too_few=0, normal_offset=0.015)
res = []
for variants in ((410, 420), (410, 420), (410, 421), (410, 421)):
    for v in variants:
        capi.check(capi.lib().smplpp_set_forward_variant(v))
    status, out = ts.step(opt, cu(theta), cu(beta), cu(vw), cu(tgt), outputs=True)
    res.append((out["e"].cpu().numpy(), out["J"].cpu().numpy()))
print("420 run-to-run: de %.3g dJ %.3g | 421 run-to-run: de %.3g dJ %.3g | 420 vs 421: de %.3g" % (
    np.abs(res[0][0] - res[1][0]).max(), np.abs(res[0][1] - res[1][1]).max(), np.abs(res[2][0] - res[3][0]).max(),
    np.abs(res[2][1] - res[3][1]).max(), np.abs(res[0][0] - res[2][0]).max()))
de = np.abs(res[0][0] - res[2][0]).reshape(F, -1)
print("420 vs 421 |de| per frame (1e-8):", np.round(de.max(1) * 1e8).astype(int))
