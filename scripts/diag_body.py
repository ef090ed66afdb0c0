import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from smplpp_b200 import api, synth
import test_ik_configs_gpu as t
f32 = np.float32
gc = dict(np.load("tests/golden/ref_ik_configs.npz"))
params = synth.make_smpl_params(0)
smpl = api.SMPL(params, device="cuda:0")
_, face_idx, _ = synth.make_marker_tasks(params)
ts = api.IkTaskSet(smpl, face_idx, vposer=api.VPoserDecoder(synth.make_vposer_params(1)))
np.set_printoptions(linewidth=250, precision=2)
n = ts.n
K = gc["body_theta"].shape[0]
th_in = np.concatenate([gc["body_theta_in"][None], gc["body_theta"][:-1]]).astype(f32)
be_in = np.concatenate([np.zeros((1, 10), f32), gc["body_beta"][:-1]]).astype(f32)
fa_in = np.concatenate([gc["face_idx"][None], gc["body_face"][:-1]]).astype(np.int32)
vw_in = np.concatenate([np.full((1, n, 3), 1.0 / 3.0, f32), gc["body_vw"][:-1]]).astype(f32)
cu = t.cu
for lo, hi, late in ((0, 25, False), (25, K, True)):
    B = hi - lo
    opt = t.body_options(api, late)
    theta, beta = cu(th_in[lo:hi]), cu(be_in[lo:hi])
    vw, face = cu(vw_in[lo:hi]), cu(fa_in[lo:hi], torch.int32)
    tgt = cu(np.repeat(gc["body_target"][None], B, axis=0))
    theta_pre, beta_pre = theta.clone(), beta.clone()
    dphi = torch.zeros((B, n, 2), device="cuda:0")
    status, out = ts.step(opt, theta, beta, vw, tgt, outputs=True, face_idx=face, dphi_out=dphi)
    print("late", late, "vw_pre diff", np.abs(vw.cpu().numpy() - gc["body_vw_pre"][lo:hi]).max())
    # point to project: compare with golden body_point
    smpl.launch(be_in[lo:hi], ts.assemble_theta(cu(th_in[lo:hi])))
    verts = smpl.getVertex()
    d = out["delta"].cpu().numpy()
    print(" |dphi| max", np.abs(dphi.cpu().numpy()).max(), "dtheta diff", np.abs(theta.cpu().numpy()-gc["body_theta"][lo:hi]).max())
    ts.reproject(opt, theta_pre, beta_pre, vw, face, dphi=dphi)
    f_gpu, f_ref = face.cpu().numpy(), gc["body_face"][lo:hi]
    agree = f_gpu == f_ref
    faces0 = smpl._faces_host.astype(np.int64) - 1
    v64 = verts.cpu().numpy().astype(np.float64)
    tri = v64[np.arange(B)[:, None, None], faces0[f_gpu]]
    pt = (vw.cpu().numpy()[..., None].astype(np.float64) * tri).sum(2)
    err = np.linalg.norm(pt - gc["body_closest"][lo:hi], axis=2)
    err[~agree] = 0
    print(" agree", agree.mean(), "closest err per iter", err.max(1))
    b, m = np.unravel_index(err.argmax(), err.shape)
    print(" worst iter", lo + b, "marker", m, "err", err[b, m], "golden point", gc["body_point"][lo + b, m], "closest", gc["body_closest"][lo + b, m], "ours", pt[b, m])
    # distance from golden point to its closest
    print(" dist", np.linalg.norm(gc["body_point"][lo+b, m] - gc["body_closest"][lo+b, m]), "tri edge lens", np.linalg.norm(tri[b, m] - np.roll(tri[b, m], 1, axis=0), axis=1))
