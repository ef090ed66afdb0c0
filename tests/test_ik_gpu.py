"""GPU parity tests of the IK step (sparse forward, analytic Jacobian, fp64 normal equations, Cholesky / box QP,
update) against golden vectors of the compiled reference harness and against the oracle."""
import numpy as np
import pytest
import torch

from conftest import TOL_JACOBIAN_REL, TOL_RESIDUAL_M, TOL_VERTEX_M

pytestmark = pytest.mark.gpu
f32 = np.float32

MODES = {
    "motion": dict(normal_task_weight=0.0, phi_limit=0.0, normal_offset=0.015, optimize_beta=0, enable_qp=1,
                   enable_phi=0),
    "body": dict(normal_task_weight=0.0, phi_limit=0.04, normal_offset=0.015, optimize_beta=1, enable_qp=1,
                 enable_phi=1),
    "interactive": dict(normal_task_weight=1.0, phi_limit=0.0, normal_offset=0.0, optimize_beta=0, enable_qp=1,
                        enable_phi=0),
    "llt": dict(normal_task_weight=0.0, phi_limit=0.0, normal_offset=0.0, optimize_beta=0, enable_qp=0, enable_phi=0),
}


def cu(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device="cuda:0").contiguous()


@pytest.fixture(scope="module")
def task_set(smpl_gpu, marker_tasks, vposer_params):
    from smplpp_b200 import api
    _, face_idx, _ = marker_tasks
    return api.IkTaskSet(smpl_gpu, face_idx, vposer=api.VPoserDecoder(vposer_params))


def test_task_positions_vs_oracle(smpl_gpu, task_set, oracle_model, marker_tasks):
    """IkTask::calcActualPos / calcActualNormal batched, with and without the 15 mm normal offset."""
    from oracle import smpl_oracle as so
    from smplpp_b200 import synth
    _, face_idx, vw = marker_tasks
    beta, theta = synth.make_forward_inputs(3, 40)
    smpl_gpu.launch(beta, theta)
    w = cu(np.repeat(vw[None], 3, axis=0))
    with torch.no_grad():
        r = so.smpl_launch(oracle_model, torch.as_tensor(beta), torch.as_tensor(theta))
    for off in (0.0, 0.015):
        # end to end: positions from the GPU forward pass (vertex tolerance)
        pos, _ = task_set.positions(smpl_gpu.getVertex(), w, off, want_normals=True)
        # kernel parity: the SAME vertices as the oracle.  A normal amplifies a vertex difference by 1 / edge length
        # (edges of the synthetic mesh go down to a few mm), so normals are only compared on identical inputs.
        pos_o, nrm_o = task_set.positions(cu(r.vertices.numpy()), w, off, want_normals=True)
        with torch.no_grad():
            for b in range(3):
                for m in (0, 7, 40):
                    t = so.IkTask(int(face_idx[m]), normal_offset=off, vertex_weights=torch.as_tensor(vw[m]))
                    ref_pos = t.calc_actual_pos(oracle_model, r.vertices[b]).numpy()
                    assert np.abs(pos[b, m].cpu().numpy() - ref_pos).max() <= TOL_VERTEX_M
                    assert np.abs(pos_o[b, m].cpu().numpy() - ref_pos).max() <= 2e-6
                    assert np.abs(nrm_o[b, m].cpu().numpy()
                                  - t.calc_actual_normal(oracle_model, r.vertices[b]).numpy()).max() < 2e-5


def test_triangle_vertex_weights_property():
    """tests/src/TestGeometryUtils.cpp:39-96 on the CUDA kernel."""
    from smplpp_b200 import api
    rng = np.random.default_rng(0)
    tri = rng.uniform(-1, 1, size=(1000, 3, 3)).astype(f32)
    w = rng.dirichlet(np.ones(3), size=1000).astype(f32)
    w[:7] = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (.5, .5, 0), (0, .5, .5), (.5, 0, .5), (1 / 3, 1 / 3, 1 / 3)]
    pos = np.einsum("ni,nik->nk", w, tri)
    got = api.calcTriangleVertexWeights(pos, tri).cpu().numpy()
    assert np.abs(got.sum(1) - 1).max() < 1e-5
    assert np.abs(np.einsum("ni,nik->nk", got, tri) - pos).max() < 1e-3


@pytest.fixture(params=[(401, 410, 420), (401, 411, 421), (402, 410, 420)], ids=["two_kernels", "two_kernels_ffma_variants", "fused_kernel"])
def ik_variant(request):
    """Both implementations of the step on a shared attachment topology: ik_jacobian_kernel + ik_solve_kernel (the
    default there) and the fused kernel (the only one for per-frame attachments)."""
    from smplpp_b200 import capi
    # 401 / 402: ik_jacobian_kernel + solve kernel / fused kernel; 410 / 411: ik_solve_mma_kernel (fp64 tensor cores, the
    # default where the problem shape allows) / the scalar ik_solve_kernel; 420 / 421: pose-blend columns of J by
    # ik_poseblend_tc_kernel (tcgen05, the default) / by the FFMA phase of ik_jacobian_kernel
    for v in request.param:
        capi.check(capi.lib().smplpp_set_forward_variant(v))
    yield request.param[0]
    for v in (400, 410, 420):
        capi.check(capi.lib().smplpp_set_forward_variant(v))


@pytest.mark.parametrize("mode", list(MODES))
def test_ik_jacobian_getter_vs_reference_golden(task_set, golden_ik, mode):
    """smplpp_ik_jacobian (the reference has no getter, callers harvest Tensor::backward rows,
    node/node.cpp:823-873): e and J against the compiled reference's rows, state untouched, weights re-seated as by step."""
    from smplpp_b200 import api
    g = golden_ik
    n = task_set.n
    opt = api.ik_options(skip_if_too_few=0, **MODES[mode])
    theta = cu(g["theta_in"].reshape(1, 75))
    beta = cu(g["beta_in"].reshape(1, 10))
    vw = cu(g["vertex_weights_in"].reshape(1, n, 3))
    tgt = cu(g[mode + "_target"].reshape(1, n, 3))
    pw = cu(g[mode + "_pos_task_weight"].reshape(1, n).astype(f32)) if (mode + "_pos_task_weight") in g else None
    theta0, beta0 = theta.clone(), beta.clone()
    e, J = task_set.jacobian(opt, theta, beta, vw, tgt, pos_task_weight=pw)
    assert torch.equal(theta, theta0) and torch.equal(beta, beta0)
    Jref = g[mode + "_J"]
    assert np.abs(e[0].cpu().numpy() - g[mode + "_e"]).max() < (2e-5 if mode == "interactive" else TOL_VERTEX_M)
    assert np.abs(J[0].cpu().numpy() - Jref).max() / np.abs(Jref).max() <= TOL_JACOBIAN_REL
    assert np.abs(vw[0].cpu().numpy() - g[mode + "_vertex_weights_out"]).max() < 1e-4
    # bitwise the rows the full step materialises
    vw2 = cu(g["vertex_weights_in"].reshape(1, n, 3))
    _, out = task_set.step(opt, theta.clone(), beta.clone(), vw2, tgt, pos_task_weight=pw, outputs=True)
    assert torch.equal(out["e"], e) and torch.equal(out["J"], J)
    e_only, none = task_set.jacobian(opt, theta, beta, cu(g["vertex_weights_in"].reshape(1, n, 3)), tgt, pos_task_weight=pw,
                                     want_jacobian=False)
    assert none is None and torch.equal(e_only, e)


@pytest.mark.parametrize("mode", list(MODES))
def test_ik_step_vs_reference_golden(task_set, golden_ik, mode, ik_variant):
    from smplpp_b200 import api
    g = golden_ik
    n = task_set.n
    opt = api.ik_options(skip_if_too_few=0, **MODES[mode])
    theta = cu(g["theta_in"].reshape(1, 75))
    beta = cu(g["beta_in"].reshape(1, 10))
    vw = cu(g["vertex_weights_in"].reshape(1, n, 3))
    tgt = cu(g[mode + "_target"].reshape(1, n, 3))
    pw = cu(g[mode + "_pos_task_weight"].reshape(1, n).astype(f32)) if (mode + "_pos_task_weight") in g else None
    status, out = task_set.step(opt, theta, beta, vw, tgt, pos_task_weight=pw, outputs=True)
    assert int(status[0]) == 0
    e, J = out["e"][0].cpu().numpy(), out["J"][0].cpu().numpy()
    Jref = g[mode + "_J"]
    assert np.abs(e - g[mode + "_e"]).max() < (2e-5 if mode == "interactive" else TOL_VERTEX_M)
    rows = np.arange(J.shape[0])
    err_pos = np.abs(J - Jref)[rows % 4 != 3].max() / np.abs(Jref).max()
    err_nrm = np.abs(J - Jref)[rows % 4 == 3].max() / np.abs(Jref).max()
    print("%s: Jacobian relative error, position rows %.3g, normal rows %.3g" % (mode, err_pos, err_nrm))
    assert np.abs(J - Jref).max() / np.abs(Jref).max() <= TOL_JACOBIAN_REL, (err_pos, err_nrm)
    # per-block relative error too (theta / phi / beta columns have different scales)
    for lo, hi in ((0, 3), (3, 75), (75, 75 + 2 * n), (75 + 2 * n, J.shape[1])):
        if hi > lo and np.abs(Jref[:, lo:hi]).max() > 0:
            assert np.abs(J[:, lo:hi] - Jref[:, lo:hi]).max() / np.abs(Jref[:, lo:hi]).max() <= 2 * TOL_JACOBIAN_REL
    bref = g[mode + "_b"]
    assert np.abs(out["b"][0].cpu().numpy() - bref).max() / np.abs(bref).max() < 1e-4
    assert np.abs(out["delta"][0].cpu().numpy() - g[mode + "_delta"]).max() < 2e-4
    assert np.abs(theta[0].cpu().numpy() - g[mode + "_theta_out"]).max() < 2e-4
    assert np.abs(beta[0].cpu().numpy() - g[mode + "_beta_out"]).max() < 2e-4
    assert np.abs(vw[0].cpu().numpy() - g[mode + "_vertex_weights_out"]).max() < 1e-4
    # A is J'J + damping: check it against the fp64 product of OUR J (bit-level definition of node.cpp:884-893)
    A = out["A"][0].cpu().numpy()
    J64 = J.astype(np.float64)
    A_chk = J64.T @ J64
    esq = float(e.astype(np.float64) @ e.astype(np.float64))
    dim = J.shape[1]
    reg = np.concatenate([np.full(75, 1e-3), np.full(2 * n, 1e-1), np.full(dim - 75 - 2 * n, 1e-3)])
    A_chk[np.diag_indices(dim)] += reg.astype(np.float32).astype(np.float64) + esq
    assert np.abs(A - A_chk).max() < 1e-9 * max(1.0, np.abs(A_chk).max())


def test_ik_step_vposer_vs_reference_golden(task_set, golden_ik, ik_variant):
    from smplpp_b200 import api
    g = golden_ik
    n = task_set.n
    opt = api.ik_options(skip_if_too_few=0, enable_vposer=1, **MODES["motion"])
    theta = cu(g["vposer_theta_in"].reshape(1, 44))
    beta = cu(g["beta_in"].reshape(1, 10))
    vw = cu(g["vertex_weights_in"].reshape(1, n, 3))
    tgt = cu(g["target_pos"].reshape(1, n, 3))
    status, out = task_set.step(opt, theta, beta, vw, tgt, outputs=True)
    assert int(status[0]) == 0
    J, Jref = out["J"][0].cpu().numpy(), g["vposer_J"]
    assert np.abs(out["e"][0].cpu().numpy() - g["vposer_e"]).max() < 5e-5
    assert np.abs(J - Jref).max() / np.abs(Jref).max() <= 2 * TOL_JACOBIAN_REL
    assert np.abs(out["delta"][0].cpu().numpy() - g["vposer_delta"]).max() < 5e-4
    assert np.abs(theta[0].cpu().numpy() - g["vposer_theta_out"]).max() < 5e-4


def test_ik_jacobian_getter_vposer_and_batch(task_set, golden_ik):
    """smplpp_ik_jacobian in VPoser mode (latent columns through the decoder Jacobian) against the compiled reference's
    rows, and on a batch of identical frames: every frame bitwise equal to the single-frame call."""
    from smplpp_b200 import api
    g = golden_ik
    n = task_set.n
    opt = api.ik_options(skip_if_too_few=0, enable_vposer=1, **MODES["motion"])
    theta = cu(g["vposer_theta_in"].reshape(1, 44))
    beta = cu(g["beta_in"].reshape(1, 10))
    vw = cu(g["vertex_weights_in"].reshape(1, n, 3))
    tgt = cu(g["target_pos"].reshape(1, n, 3))
    e, J = task_set.jacobian(opt, theta, beta, vw, tgt)
    Jref = g["vposer_J"]
    assert np.abs(e[0].cpu().numpy() - g["vposer_e"]).max() < 5e-5
    assert np.abs(J[0].cpu().numpy() - Jref).max() / np.abs(Jref).max() <= 2 * TOL_JACOBIAN_REL
    B = 37
    thB = theta.repeat(B, 1).contiguous()
    vwB = cu(np.repeat(g["vertex_weights_in"].reshape(1, n, 3), B, axis=0))
    tgB = tgt.repeat(B, 1, 1).contiguous()
    eB, JB = task_set.jacobian(opt, thB, beta.repeat(B, 1).contiguous(), vwB, tgB)
    assert torch.equal(eB, e.expand(B, -1)) and torch.equal(JB, J.expand(B, -1, -1))
    assert torch.equal(vwB, vw.expand(B, -1, -1))


def test_ik_converges_like_oracle(task_set, oracle_model, marker_tasks, smpl_gpu):
    """Run K iterations of the motion-mode step on 3 frames and compare the marker residual trajectory with the
    oracle's (converged residual within 1e-4 m)."""
    from oracle import smpl_oracle as so
    from smplpp_b200 import api, synth
    _, face_idx, vw0 = marker_tasks
    n, B, K = task_set.n, 2, 6
    gt = synth.make_motion(40, 20)[[5, 30]]
    beta = (np.random.default_rng(5).normal(size=10) * 0.5).astype(f32)
    smpl_gpu.launch(beta, gt)
    w0 = cu(np.repeat(vw0[None], B, axis=0))
    target = task_set.positions(smpl_gpu.getVertex(), w0, 0.015).contiguous()
    x0 = np.repeat(synth.make_motion(40, 20)[:1].reshape(1, 75), B, axis=0)
    opt = api.ik_options(**MODES["motion"])
    theta, vw = cu(x0), w0.clone()
    beta_d = cu(beta)
    res_gpu = []
    for _ in range(K):
        status, out = task_set.step(opt, theta, beta_d, vw, target, outputs=True)
        assert (status == 0).all()
        res_gpu.append(np.linalg.norm(out["e"].cpu().numpy().reshape(B, n, 4)[:, :, :3], axis=2).mean(axis=1))
    tgt_h = target.cpu().numpy()
    for b in range(B):
        tasks = [so.IkTask(int(face_idx[i]), target_pos=torch.as_tensor(tgt_h[b, i]), normal_task_weight=0.0,
                           phi_limit=0.0, normal_offset=0.015, vertex_weights=torch.as_tensor(vw0[i])) for i in range(n)]
        x = x0[b].copy()
        for k in range(K):
            r = so.ik_iteration(oracle_model, tasks, x, beta, skip_if_too_few=True)
            res_o = np.linalg.norm(r.e.reshape(n, 4)[:, :3], axis=1).mean()
            assert abs(res_o - res_gpu[k][b]) < TOL_RESIDUAL_M
            x = r.theta_state
        assert np.abs(theta[b].cpu().numpy() - x).max() < 5e-3
    assert res_gpu[-1].max() < res_gpu[0].min()


def test_ik_skip_and_missing_markers(task_set, marker_tasks, smpl_gpu):
    """node.cpp:785: a motion-mode frame with fewer than n/2 valid markers is left untouched (status 1)."""
    from smplpp_b200 import api, synth
    n = task_set.n
    x0 = synth.initial_theta(False).reshape(1, 75).repeat(2, axis=0)
    theta = cu(x0)
    beta = cu(np.zeros(10, f32))
    vw = task_set.default_vertex_weights(2)
    tgt = cu(np.random.default_rng(1).normal(size=(2, n, 3)).astype(f32) * 0.3)
    pw = np.ones((2, n), f32)
    pw[1, : n - 10] = 0.0
    status = task_set.step(api.ik_options(**MODES["motion"]), theta, beta, vw, tgt, pos_task_weight=cu(pw))
    assert status.cpu().tolist() == [0, 1]
    assert np.array_equal(theta[1].cpu().numpy(), x0[1])
    assert not np.array_equal(theta[0].cpu().numpy(), x0[0])


def test_shared_beta_single_frame_equals_joint_qp(task_set, oracle_model, marker_tasks, golden_ik, ik_variant):
    """With ONE frame the shared-beta stage (Schur complement + 10-dim box QP + back-substitution) must equal
    the reference's joint QP over [theta | beta] with |dbeta| <= 0.5 (phi pinned)."""
    from oracle import smpl_oracle as so
    from smplpp_b200 import api
    g = golden_ik
    _, face_idx, _ = marker_tasks
    n = task_set.n
    tasks = [so.IkTask(int(face_idx[i]), target_pos=torch.as_tensor(g["target_pos"][i]), normal_task_weight=0.0,
                       phi_limit=0.0, normal_offset=0.015, vertex_weights=torch.as_tensor(g["vertex_weights_in"][i]))
             for i in range(n)]
    r = so.ik_iteration(oracle_model, tasks, g["theta_in"], g["beta_in"], optimize_beta=True)
    opt = api.ik_options(skip_if_too_few=0, **MODES["motion"])
    theta = cu(g["theta_in"].reshape(1, 75))
    sbeta = cu(g["beta_in"].reshape(10))
    vw = cu(g["vertex_weights_in"].reshape(1, n, 3))
    status = task_set.shared_beta_step(opt, theta, sbeta, vw, cu(g["target_pos"].reshape(1, n, 3)))
    assert int(status[0]) == 0
    assert np.abs(sbeta.cpu().numpy() - r.beta).max() < 2e-4
    assert np.abs(theta[0].cpu().numpy() - r.theta_state).max() < 2e-4


def test_shared_beta_many_frames_kkt(task_set, marker_tasks, smpl_gpu, ik_variant):
    """Many frames: the reduced 111 doubles are a deterministic sum, the solve satisfies the box-QP KKT
    conditions, and two identical runs agree bitwise."""
    from smplpp_b200 import api, synth
    _, face_idx, vw0 = marker_tasks
    n, B = task_set.n, 37
    gt = synth.make_motion(B, 23)
    beta_true = (np.random.default_rng(6).normal(size=10)).astype(f32)
    smpl_gpu.launch(beta_true, gt)
    w0 = cu(np.repeat(vw0[None], B, axis=0))
    target = task_set.positions(smpl_gpu.getVertex(), w0, 0.015).contiguous()
    opt = api.ik_options(**MODES["motion"])
    outs = []
    for _ in range(2):
        theta = cu(gt.reshape(B, 75))
        sbeta = cu(np.zeros(10, f32))
        vw = w0.clone()
        status, red = task_set.shared_beta_step(opt, theta, sbeta, vw, target, return_reduced=True)
        assert (status == 0).all()
        outs.append((theta.cpu().numpy(), sbeta.cpu().numpy(), red.cpu().numpy()))
    assert all(np.array_equal(a, b) for a, b in zip(outs[0], outs[1]))
    red = outs[0][2]
    S = red[:100].reshape(10, 10) + (1e-3 + red[110]) * np.eye(10)
    x = outs[0][1].astype(np.float64)  # beta started at 0 => beta == dbeta
    grad = S @ x + red[100:110]
    assert np.abs(S - S.T).max() < 1e-9 * np.abs(S).max()
    free = np.abs(x) < 0.5 - 1e-6
    assert np.abs(grad[free]).max(initial=0.0) < 1e-5 * max(1.0, np.abs(red[100:110]).max())
    assert (grad[x >= 0.5 - 1e-6] <= 1e-6).all() and (grad[x <= -0.5 + 1e-6] >= -1e-6).all()
    # moving towards the true shape
    assert np.linalg.norm(x - beta_true) < np.linalg.norm(beta_true)


@pytest.mark.parametrize("n_tasks,seed", [(77, 1), (5, 2)])
def test_tensor_core_stages_equal_ffma_on_random_task_sets(smpl_gpu, params, n_tasks, seed):
    """The tensor-core stages of the two-kernel path (pose-blend columns on tcgen05, normal equations on the fp64 tensor
    cores) against their FFMA / scalar predecessors on task sets other than the 41 markers of the goldens: random faces,
    37 frames (one full and one partial frame block), with and without normal rows."""
    from smplpp_b200 import api, capi, synth
    rng = np.random.default_rng(seed)
    faces = rng.choice(params.face_indices.shape[0], size=n_tasks, replace=False).astype(np.int64)
    ts = api.IkTaskSet(smpl_gpu, faces)
    F = 37
    theta = synth.make_motion(F, 7 + seed).reshape(F, 75).astype(f32)
    theta[:, 3:] += rng.normal(scale=0.2, size=(F, 72)).astype(f32)
    beta = (rng.normal(size=10) * 0.5).astype(f32)
    vw = rng.dirichlet(np.ones(3), size=(F, n_tasks)).astype(f32)
    tgt = rng.normal(scale=0.5, size=(F, n_tasks, 3)).astype(f32)
    tn = rng.normal(size=(F, n_tasks, 3)).astype(f32)
    tn /= np.linalg.norm(tn, axis=2, keepdims=True)
    for kw in (dict(normal_offset=0.015), dict(normal_task_weight=1.0, normal_offset=0.0), dict(normal_offset=0.0)):
        opt = api.ik_options(update_state=0, skip_if_too_few=0, **kw)
        outs = {}
        for variants in ((411, 421), (410, 422), (410, 420)):
            for v in variants:
                capi.check(capi.lib().smplpp_set_forward_variant(v))
            status, out = ts.step(opt, cu(theta), cu(beta), cu(vw), cu(tgt),
                                  target_normal=cu(tn) if kw.get("normal_task_weight") else None, outputs=True)
            outs[variants] = (status.cpu().numpy(), out["J"].cpu().numpy(), out["delta"].cpu().numpy(), out["A"].cpu().numpy())
        for v in (410, 420):
            capi.check(capi.lib().smplpp_set_forward_variant(v))
        (s0, J0, d0, A0), (s1, J1, d1, A1), (s2, J2, d2, A2) = outs[(411, 421)], outs[(410, 422)], outs[(410, 420)]
        assert np.array_equal(s0, s1) and np.array_equal(s0, s2) and (s0 == 0).all()
        rel = lambda a, b: np.abs(a - b).reshape(F, -1).max(1) / np.abs(b).reshape(F, -1).max(1)
        # same rest shape (FFMA), tensor-core stages against their predecessors: rounding level in every frame
        assert rel(J1, J0).max() < 2e-6
        assert rel(A1, A0).max() < 1e-5 and np.abs(d1 - d0).max() < 1e-5
        # the default path also takes the rest shape from tcgen05: it differs from the FFMA one in the last bit (3e-8 m,
        # test_task_rest_shape_kernels).  Where the face normals around a corner nearly cancel (random poses fold the
        # synthetic mesh) the vertex normal amplifies that a hundred- to a thousandfold (scripts/diag_rest.py), so this
        # comparison is a sanity bound, not a parity statement
        assert np.median(rel(J2, J0)) < 1e-4 and rel(J2, J0).max() < 2e-2


def test_task_rest_shape_kernels(smpl_gpu, params, task_set):
    """smplpp_tasks_rest_shape: the rest shape of the task vertices on tcgen05 (what the IK step runs per iteration) against
    the FFMA kernel and against SMPL::getRestShape of the full model at those vertices."""
    from smplpp_b200 import synth
    F = 150  # one full 128-frame tile and a partial one
    rng = np.random.default_rng(4)
    theta = synth.make_motion(F, 9).reshape(F, 75).astype(f32)
    theta[:, 3:] += rng.normal(scale=0.2, size=(F, 72)).astype(f32)
    beta = (rng.normal(size=(F, 10)) * 0.7).astype(f32)
    r_tc, ids = task_set.restShape(beta, theta, 0)
    r_ff, ids2 = task_set.restShape(beta, theta, 1)
    assert np.array_equal(ids, ids2) and len(ids) == task_set.vertex_count
    smpl_gpu.launch(beta, theta.reshape(F, 25, 3))
    full = smpl_gpu.getRestShape()[:, torch.as_tensor(ids.astype(np.int64), device="cuda:0")]
    assert (r_ff - full).abs().max().item() < 2e-7
    assert (r_tc - full).abs().max().item() < 2e-7
