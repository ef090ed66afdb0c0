import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smplpp_b200 import api, capi, synth
params = synth.make_smpl_params(0)
smpl = api.SMPL(params, device="cuda:0")
for n_tasks, seed in ((5, 2), (77, 1), (41, 0)):
    rng = np.random.default_rng(seed)
    faces = rng.choice(params.face_indices.shape[0], size=n_tasks, replace=False).astype(np.int64)
    ts = api.IkTaskSet(smpl, faces)
    F = 150
    theta = synth.make_motion(F, 7 + seed).reshape(F, 75).astype(np.float32)
    theta[:, 3:] += rng.normal(scale=0.2, size=(F, 72)).astype(np.float32)
    beta = (rng.normal(size=10) * 0.5).astype(np.float32)
    r_tc, ids = ts.restShape(beta, theta, 0)
    r_ff, _ = ts.restShape(beta, theta, 1)
    smpl.launch(beta, theta.reshape(F, 25, 3))
    full = smpl.getRestShape()[:, torch.as_tensor(ids.astype(np.int64), device="cuda:0")]
    # float64 numpy
    from oracle import smpl_oracle as so
    print(n_tasks, "nU", len(ids), "tc vs ffma %.3g  tc vs full-model rest %.3g  ffma vs full %.3g  max|rest| %.3g" % (
        (r_tc - r_ff).abs().max().item(), (r_tc - full).abs().max().item(), (r_ff - full).abs().max().item(), full.abs().max().item()))
    d = (r_tc - r_ff).abs().cpu().numpy()
    idx = np.unravel_index(d.argmax(), d.shape); print("   worst (frame, vertex, axis)", idx, "per-frame max", np.round(d.reshape(F, -1).max(1)[:40] * 1e7) / 10)
