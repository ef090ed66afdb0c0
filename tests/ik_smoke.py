"""smoke() leg for the IK step (test infrastructure, called by __graft_entry__.smoke()): one frame, one iteration on
cuda:0 checked against the oracle."""
import numpy as np
import torch

from smplpp_b200 import api, synth


def run(params):
    from oracle import smpl_oracle as so
    smpl = api.SMPL(params, device="cuda:0")
    _, face_idx, vw = synth.make_marker_tasks(params)
    tasks = api.IkTaskSet(smpl, face_idx)
    n = tasks.n
    gt = synth.make_motion(4, 20)
    beta = np.zeros(10, np.float32)
    smpl.launch(beta, gt[3:4])
    w = torch.as_tensor(vw[None], device="cuda:0").contiguous()
    target = tasks.positions(smpl.getVertex(), w, 0.015).contiguous()
    theta = torch.as_tensor(gt[:1].reshape(1, 75), device="cuda:0").contiguous()
    opt = api.ik_options()
    status, out = tasks.step(opt, theta, torch.as_tensor(beta, device="cuda:0"), w.clone(), target, outputs=True)
    model = so.SmplModel.from_params(params)
    tgt = target[0].cpu().numpy()
    otasks = [so.IkTask(int(face_idx[i]), target_pos=torch.as_tensor(tgt[i]), normal_task_weight=0.0, phi_limit=0.0,
                        normal_offset=0.015, vertex_weights=torch.as_tensor(vw[i])) for i in range(n)]
    r = so.ik_iteration(model, otasks, gt[0].reshape(-1), beta)
    J = out["J"][0].cpu().numpy()
    rel = float(np.abs(J - r.J).max() / np.abs(r.J).max())
    dth = float(np.abs(theta[0].cpu().numpy() - r.theta_state).max())
    print("smoke: IK step (41 markers, D=75) Jacobian rel err = %.3g (tol 1e-4), |dtheta| = %.3g" % (rel, dth))
    assert int(status[0]) == 0 and rel <= 1e-4 and dth < 2e-4
