"""CPU suite: pins the oracle (oracle/smpl_oracle.py) against the reference's own known answers and against
golden vectors produced by the compiled reference (tests/golden/make_ref_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import smpl_oracle as so

f32 = np.float32


def t(a):
    return torch.as_tensor(np.asarray(a, dtype=f32))


# ---- known-answer vectors of src/toolbox/Tester.cpp (printed there to 6 decimals) ----

def test_kat_blend_shape(kat):
    i, e = kat["blendShape"]["inputs"], kat["blendShape"]["expected"]
    rot = so.rodrigues(t(i["theta"]))
    pbs = so.pose_blend(so.pose_blend_coeffs(rot), t(i["poseBlendBasis"]))
    sbs = so.shape_blend(t(i["beta"]), t(i["shapeBlendBasis"]))
    assert np.allclose(sbs.numpy().reshape(-1), np.asarray(e["shapeBlendShape"]).reshape(-1), atol=2e-6)
    assert np.allclose(pbs.numpy().reshape(-1), np.asarray(e["poseBlendShape"]).reshape(-1), atol=5e-6)
    assert np.allclose(rot.numpy()[0, :5], np.asarray(e["poseRotation"]), atol=1.5e-6)


def test_kat_joint_regression(kat):
    i, e = kat["jointRegression"]["inputs"], kat["jointRegression"]["expected"]
    rest = so.linear_combine(t(i["templateShape"]), t(i["shapeBlendShape"]), t(i["poseBlendShape"]))
    joints = so.joint_regress(t(i["templateShape"]), t(i["shapeBlendShape"]), t(i["jointRegressor"]))
    assert np.allclose(rest.numpy(), np.asarray(e["restShape"]), atol=1.5e-6)
    assert np.allclose(joints.numpy()[0], np.asarray(e["joints"]), atol=2e-6)


def test_kat_world_transformation(kat):
    i, e = kat["worldTransformation"]["inputs"], kat["worldTransformation"]["expected"]
    parents = [int(p) if p < 24 else -1 for p in i["kineTree"][0]]
    rel, _ = so.world_transform(t(i["poseRotation"]), t(i["joints"]), parents)
    exp = np.asarray(e["transformations"])
    assert np.allclose(rel.numpy()[0, :5], exp, atol=2e-6, rtol=2e-6)


def test_kat_linear_blend_skinning(kat):
    i, e = kat["linearBlendSkinning"]["inputs"], kat["linearBlendSkinning"]["expected"]
    v = so.skinning(t(i["weights"]), t(i["restShape"]), t(i["transformations"]))
    assert np.allclose(v.numpy().reshape(-1), np.asarray(e["vertices"]).reshape(-1), atol=1.5e-6)


# ---- golden vectors from the compiled reference ----

def test_forward_vs_reference_golden(oracle_model, golden_forward):
    g = golden_forward
    v, j, rest, _ = so.forward_numpy(oracle_model, g["beta"], g["theta"])
    assert np.abs(v - g["vertices"]).max() < 2e-6
    assert np.abs(j - g["joints"]).max() < 1e-6
    assert np.abs(rest - g["rest_shape"]).max() < 1e-6


def test_normals_vs_reference_golden(oracle_model, golden_forward):
    g = golden_forward
    with torch.no_grad():
        verts0 = so.smpl_launch(oracle_model, t(g["beta"][:1]), t(g["theta"][:1])).vertices[0]
        fn = np.stack([so.calc_normal(oracle_model, verts0, int(f)).numpy() for f in g["normal_face_idx"]])
        vn = np.stack([so.calc_vertex_normal(oracle_model, verts0, int(v)).numpy() for v in g["normal_vert_idx"][:20]])
    assert np.abs(fn - g["face_normals"]).max() < 5e-5
    assert np.abs(vn - g["vertex_normals"][:20]).max() < 5e-5


def test_vposer_vs_reference_golden(oracle_vposer, golden_vposer):
    g = golden_vposer
    with torch.no_grad():
        aa = oracle_vposer.forward(t(g["latent"])).numpy()
    assert np.abs(aa - g["axis_angle"]).max() < 2e-5
    z = t(g["latent"][:1]).requires_grad_(True)
    out = oracle_vposer.forward(z).view(-1)
    jac = np.stack([torch.autograd.grad(out[r], z, retain_graph=True)[0].numpy()[0] for r in range(63)])
    ref = g["jacobian"][0]
    assert np.abs(jac - ref).max() / np.abs(ref).max() < 1e-4


def test_rotmat_to_axis_angle_vs_reference_golden(golden_vposer):
    g = golden_vposer
    r = t(g["prop_rotmat"]).requires_grad_(True)
    aa = so.rotmat_to_axis_angle(r)
    aa.sum().backward()
    assert np.abs(aa.detach().numpy() - g["prop_axis_angle"]).max() < 1e-6
    assert np.isfinite(r.grad.numpy()).all()
    scale = np.abs(g["prop_grad"]).max()
    assert np.abs(r.grad.numpy() - g["prop_grad"]).max() / scale < 1e-5


@pytest.mark.parametrize("mode", ["motion", "body", "interactive", "llt"])
def test_ik_iteration_vs_reference_golden(oracle_model, golden_ik, mode):
    g = golden_ik
    n = len(g["face_idx"])
    kw = dict(motion=dict(nw=0.0, pl=0.0, no=0.015, ob=False, qp=True),
              body=dict(nw=0.0, pl=0.04, no=0.015, ob=True, qp=True),
              interactive=dict(nw=1.0, pl=0.0, no=0.0, ob=False, qp=True),
              llt=dict(nw=0.0, pl=0.0, no=0.0, ob=False, qp=False))[mode]
    pw = g.get(mode + "_pos_task_weight", np.ones(n))
    tasks = [so.IkTask(int(g["face_idx"][i]), target_pos=t(g[mode + "_target"][i]), pos_task_weight=float(pw[i]),
                       normal_task_weight=kw["nw"], phi_limit=kw["pl"], normal_offset=kw["no"],
                       vertex_weights=t(g["vertex_weights_in"][i])) for i in range(n)]
    r = so.ik_iteration(oracle_model, tasks, g["theta_in"], g["beta_in"], optimize_beta=kw["ob"], enable_qp=kw["qp"])
    J = g[mode + "_J"].astype(np.float64)
    # rows of the normal task sum vertex normals in the (unordered_map) order of SMPL.cpp:529: 1e-5 noise
    loose = 10.0 if mode == "interactive" else 1.0
    assert np.abs(r.e - g[mode + "_e"]).max() < 2e-6 * loose
    assert np.abs(r.J - J).max() / np.abs(J).max() < 2e-5 * loose
    assert np.abs(r.b - g[mode + "_b"]).max() / np.abs(g[mode + "_b"]).max() < 2e-5 * loose
    assert np.abs(r.delta - g[mode + "_delta"]).max() < 2e-6 * loose
    assert np.abs(r.theta_state - g[mode + "_theta_out"]).max() < 2e-6 * loose
    assert np.abs(r.beta - g[mode + "_beta_out"]).max() < 2e-6
    assert np.abs(r.vertex_weights - g[mode + "_vertex_weights_out"]).max() < 1e-5


def test_ik_iteration_vposer_vs_reference_golden(oracle_model, oracle_vposer, golden_ik):
    g = golden_ik
    n = len(g["face_idx"])
    tasks = [so.IkTask(int(g["face_idx"][i]), target_pos=t(g["target_pos"][i]), normal_task_weight=0.0, phi_limit=0.0,
                       normal_offset=0.015, vertex_weights=t(g["vertex_weights_in"][i])) for i in range(n)]
    r = so.ik_iteration(oracle_model, tasks, g["vposer_theta_in"], g["beta_in"], vposer=oracle_vposer)
    J = g["vposer_J"].astype(np.float64)
    assert np.abs(r.e - g["vposer_e"]).max() < 5e-6
    assert np.abs(r.J - J).max() / np.abs(J).max() < 5e-5
    assert np.abs(r.delta - g["vposer_delta"]).max() < 1e-5
    assert np.abs(r.theta_state - g["vposer_theta_out"]).max() < 1e-5


# ---- the reference's own property tests, run on the oracle ----

def test_triangle_vertex_weights_property():
    """tests/src/TestGeometryUtils.cpp:39-96: weights sum to 1 and reproduce the point (1e-3)."""
    rng = np.random.default_rng(0)
    corner = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (.5, .5, 0), (0, .5, .5), (.5, 0, .5), (1 / 3, 1 / 3, 1 / 3)]
    for _ in range(100):
        tri = rng.uniform(-1, 1, size=(3, 3)).astype(f32)
        ws = corner + [tuple(w) for w in rng.dirichlet(np.ones(3), size=10)]
        for w in ws:
            w = np.asarray(w, dtype=f32)
            pos = w @ tri
            got = so.triangle_vertex_weights(t(pos), t(tri)).numpy()
            assert abs(got.sum() - 1) < 1e-5
            assert np.abs(got @ tri - pos).max() < 1e-3


def test_rotmat_to_axis_angle_property(golden_vposer):
    """tests/src/TestVPoser.cpp:16-70: ||aa - truth|| < 5e-3 (sign flip allowed at theta ~ pi), NaN-free grads."""
    from scipy.spatial.transform import Rotation
    g = golden_vposer
    r = t(g["prop_rotmat"]).requires_grad_(True)
    aa = so.rotmat_to_axis_angle(r)
    aa.sum().backward()
    got = aa.detach().numpy().astype(np.float64)
    truth = Rotation.from_matrix(g["prop_rotmat"].astype(np.float64)).as_rotvec()
    err = np.minimum(np.linalg.norm(got - truth, axis=1), np.linalg.norm(got + truth, axis=1))
    # the last 60 inputs (random axes within 1e-2 of pi) are outside the reference test's families: with the
    # (1 - eps) bias inside arccos the sign rule of VPoser.cpp:98-101 acts on rounding noise there.  They are
    # covered by the golden comparison above; here only NaN-freeness is asserted for them.
    fam = np.arange(err.shape[0]) < err.shape[0] - 60
    assert (err[fam] < 5e-3).all()
    assert np.isfinite(got).all() and np.isfinite(r.grad.numpy()).all()


def test_box_qp_kkt():
    """The QP backend is third-party in the reference (parity unpinned): check KKT optimality instead."""
    rng = np.random.default_rng(4)
    for n in (5, 30, 90):
        M = rng.normal(size=(n + 5, n))
        A = M.T @ M + 1e-3 * np.eye(n)
        b = rng.normal(size=n) * 3
        lo = np.where(rng.random(n) < 0.5, -0.04, -np.inf)
        hi = -lo
        lo[:3] = hi[:3] = 0.0
        x = so.solve_box_qp(A, b, lo, hi)
        g = A @ x + b
        assert (x >= lo - 1e-12).all() and (x <= hi + 1e-12).all()
        free = (x > lo + 1e-12) & (x < hi - 1e-12)
        assert np.abs(g[free]).max(initial=0) < 1e-8
        at_lo = (np.abs(x - lo) <= 1e-12) & (lo != hi)
        at_hi = (np.abs(x - hi) <= 1e-12) & (lo != hi)
        assert (g[at_lo] >= -1e-8).all() and (g[at_hi] <= 1e-8).all()


# ---- projection onto the mesh (node/node.cpp:970-1001; libigl is un-vendored: property checks pin the restatement) ----

def test_oracle_mesh_projection_properties(params):
    """The restated point-mesh projection: (1) a point lifted off a face along its normal by less than the local
    feature size projects back to its foot with the barycentric weights it was built from; (2) no sampled point of
    any face is closer than the reported minimum; (3) a point at a vertex returns that vertex, the lowest incident
    face index, and weights that reproduce it."""
    model = so.SmplModel.from_params(params)
    faces0 = (model.face_indices.numpy() - 1)
    verts = np.asarray(params.vertices_template, dtype=np.float64)
    rng = np.random.default_rng(5)
    fidx = rng.choice(len(faces0), size=24, replace=False)
    bary = rng.dirichlet(np.ones(3) * 2.0, size=24)
    tri = verts[faces0[fidx]]                                  # (24,3,3)
    foot = (bary[:, :, None] * tri).sum(1)
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    pts = foot + 1e-4 * nrm                                    # 0.1 mm above the face
    face, closest, sq, w = so.project_points_on_mesh(verts, faces0, pts)
    # (2) brute-force sampling of EVERY face on a barycentric grid never beats the reported distance
    g = np.array([[i, j, 8 - i - j] for i in range(9) for j in range(9 - i)], dtype=np.float64) / 8.0
    samples = np.einsum("gk,fkd->fgd", g, verts[faces0])      # (F, 45, 3)
    for i in range(len(pts)):
        d2 = ((samples - pts[i]) ** 2).sum(-1).min()
        assert sq[i] <= d2 + 1e-15
        # the reported closest point lies on the reported face and reproduces the distance
        t = verts[faces0[face[i]]]
        assert np.abs((w[i][:, None] * t).sum(0) - closest[i]).max() < 1e-9
        assert abs(((closest[i] - pts[i]) ** 2).sum() - sq[i]) < 1e-15
    # (1) convex-ish neighbourhoods: whenever the own face is the answer the foot and the weights come back exactly
    own = face == fidx
    assert own.sum() >= 12
    assert np.abs(closest[own] - foot[own]).max() < 1e-9
    assert np.abs(w[own] - bary[own]).max() < 1e-6
    assert np.all(sq <= 1e-8 + 1e-12)                          # never farther than the 0.1 mm lift
    # (3) exactly at a vertex
    vid = int(faces0[fidx[0], 1])
    face_v, closest_v, sq_v, w_v = so.project_points_on_mesh(verts, faces0, verts[vid][None])
    incident = np.nonzero((faces0 == vid).any(1))[0]
    assert face_v[0] == incident.min() and sq_v[0] == 0.0
    assert np.abs(closest_v[0] - verts[vid]).max() == 0.0
    assert np.abs((w_v[0][:, None] * verts[faces0[face_v[0]]]).sum(0) - verts[vid]).max() < 1e-12


def test_sweep_grid_oracle_on_a_cube():
    """Winding number of a closed cube: 1 inside, 0 outside, 1/2 on a face, and the grid bounds of GridUtils.hpp."""
    from oracle import smpl_oracle as so
    c = np.array([[x, y, z] for x in (0.0, 1.0) for y in (0.0, 1.0) for z in (0.0, 1.0)]) * 0.1 + np.array([0.013, -0.06, 0.2])
    faces = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4],
                      [1, 5, 7], [1, 7, 3]])
    lo, num, w = so.sweep_grid(c, faces)
    assert np.array_equal(lo, [0, -3, 8]) and np.array_equal(lo + num - 1, [5, 2, 12])
    pts = 0.025 * (np.stack(np.meshgrid(*[np.arange(n) for n in num], indexing="ij"), -1) + lo)
    inside = ((pts > c.min(0) + 1e-6) & (pts < c.max(0) - 1e-6)).all(-1)
    outside = ((pts < c.min(0) - 1e-6) | (pts > c.max(0) + 1e-6)).any(-1)
    sign = 1.0 if w[inside].mean() > 0 else -1.0  # orientation of the hand-written faces
    assert inside.sum() >= 27 and np.abs(sign * w[inside] - 1.0).max() < 1e-9
    assert np.abs(w[outside]).max() < 1e-9
