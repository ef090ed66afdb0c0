"""GPU parity of smplpp_closest_points (node/node.cpp:970-1001: projection of the task points onto the posed mesh and
the re-seated face / vertex weights) against the float64 oracle restatement, through the C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_POINT_M = 2e-6   # fp32 closest point on a mesh with ~1 m coordinates vs the float64 oracle


def _oracle(verts, faces0, pts):
    from oracle import smpl_oracle as so
    return so.project_points_on_mesh(verts, faces0, pts)


def _check_frame(verts, faces0, pts, face, closest, sq, w):
    o_face, o_closest, o_sq, o_w = _oracle(verts, faces0, pts)
    v64 = verts.astype(np.float64)
    for i in range(len(pts)):
        # same distance; same closest point (the minimiser is unique unless two faces tie, in which case the point is shared)
        assert abs(np.sqrt(sq[i]) - np.sqrt(o_sq[i])) <= TOL_POINT_M
        if face[i] != o_face[i]:
            # a tie (shared edge / vertex) or two faces within fp32 resolution: the oracle's distance to OUR face must match
            from oracle import smpl_oracle as so
            t = v64[faces0[face[i]]]
            q = so.closest_points_on_triangles(pts[i].astype(np.float64), t[0:1], t[1:2], t[2:3])[0]
            assert abs(np.linalg.norm(q - pts[i]) - np.sqrt(o_sq[i])) <= TOL_POINT_M
        else:
            assert np.abs(closest[i] - o_closest[i]).max() <= TOL_POINT_M
        # the re-seated weights are a convex combination that reproduces the closest point on the returned face
        assert abs(w[i].sum() - 1.0) < 1e-5 and w[i].min() >= 0.0
        assert np.abs((w[i][:, None] * v64[faces0[face[i]]]).sum(0) - closest[i]).max() <= 5e-6
    return o_face


def test_closest_points_vs_oracle(smpl_gpu, oracle_model, marker_tasks, params):
    from smplpp_b200 import synth
    faces0 = (np.asarray(params.face_indices, dtype=np.int64) - 1)
    _, face_idx, vw = marker_tasks
    B = 3
    beta, theta = synth.make_forward_inputs(B, 91)
    smpl_gpu.launch(beta, theta)
    verts = smpl_gpu.getVertex()
    vh = verts.cpu().numpy()
    rng = np.random.default_rng(3)
    # marker-like points: on the attachment faces, lifted 15 mm along the face normal, plus 5 mm of noise
    tri = vh[:, faces0[np.asarray(face_idx)]]                                  # (B, n, 3, 3)
    foot = (np.asarray(vw)[None, :, :, None] * tri).sum(2)
    nrm = np.cross(tri[:, :, 1] - tri[:, :, 0], tri[:, :, 2] - tri[:, :, 0])
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    pts = (foot + 0.015 * nrm + rng.normal(0, 0.005, foot.shape)).astype(np.float32)
    face, closest, sq, w = smpl_gpu.projectPoints(pts)
    face, closest, sq, w = face.cpu().numpy(), closest.cpu().numpy(), sq.cpu().numpy(), w.cpu().numpy()
    same = 0
    for b in range(B):
        o_face = _check_frame(vh[b], faces0, pts[b], face[b], closest[b], sq[b], w[b])
        same += int((o_face == face[b]).sum())
    assert same >= 0.95 * face.size  # ties are rare


def test_closest_points_edge_cases(smpl_gpu, params):
    """Points exactly at vertices (ties between the incident faces resolve to the lowest face index in both
    implementations), far outside the body, one point per frame, 256 points per frame, ragged batch."""
    from smplpp_b200 import capi, synth
    faces0 = (np.asarray(params.face_indices, dtype=np.int64) - 1)
    beta, theta = synth.make_forward_inputs(2, 17)
    smpl_gpu.launch(beta, theta)
    vh = smpl_gpu.getVertex().cpu().numpy()
    rng = np.random.default_rng(9)
    vid = rng.choice(vh.shape[1], size=256, replace=False)
    pts = np.stack([vh[0, vid], vh[1, vid]]).astype(np.float32)
    pts[:, 200:] += rng.normal(0, 3.0, (2, 56, 3)).astype(np.float32)          # far away
    face, closest, sq, w = [t.cpu().numpy() for t in smpl_gpu.projectPoints(pts)]
    for b in range(2):
        o_face = _check_frame(vh[b], faces0, pts[b], face[b], closest[b], sq[b], w[b])
        at_vertex = np.arange(200)
        assert np.all(sq[b][at_vertex] == 0.0)
        assert np.array_equal(face[b][at_vertex], o_face[at_vertex])           # lowest incident face
        assert np.abs(closest[b][at_vertex] - pts[b][at_vertex]).max() == 0.0
    # one point per frame
    f1, c1, s1, w1 = [t.cpu().numpy() for t in smpl_gpu.projectPoints(pts[:, 7:8])]
    assert np.array_equal(f1[:, 0], face[:, 7]) and np.array_equal(c1[:, 0], closest[:, 7])
    # more than 512 points per frame is refused with the reference-style message
    with pytest.raises(capi.SmplppError, match="Failed to project points onto the mesh"):
        smpl_gpu.projectPoints(np.zeros((2, 513, 3), np.float32))


def test_closest_points_round_trip_full_batch(smpl_gpu, marker_tasks, params):
    """Size-independent property at a mocap-sized batch: points built ON the attachment faces project to distance 0
    (to fp32 resolution) and the re-seated weights reproduce them."""
    from smplpp_b200 import synth
    faces0 = torch.as_tensor(np.asarray(params.face_indices, dtype=np.int64) - 1, device="cuda:0")
    _, face_idx, vw = marker_tasks
    B = 1024
    beta, theta = synth.make_forward_inputs(B, 23)
    smpl_gpu.launch(beta, theta)
    verts = smpl_gpu.getVertex()
    tri = verts[:, faces0[torch.as_tensor(np.asarray(face_idx), device="cuda:0")]]   # (B, n, 3, 3)
    wt = torch.as_tensor(np.asarray(vw, dtype=np.float32), device="cuda:0")
    pts = (wt[None, :, :, None] * tri).sum(2).contiguous()
    face, closest, sq, w = smpl_gpu.projectPoints(pts)
    assert float(sq.max()) < (2e-6) ** 2
    tri2 = verts[torch.arange(B, device="cuda:0")[:, None, None], faces0[face.long()]]   # (B, n, 3, 3)
    rec = (w[..., None] * tri2).sum(2)
    assert float((rec - pts).abs().max()) < 5e-6


@pytest.mark.parametrize("with_phi", [False, True])
def test_reproject_vs_oracle(smpl_gpu, oracle_model, marker_tasks, params, vposer_params, with_phi):
    """IkTaskSet.reproject = the tail of the reference's IK iteration (node/node.cpp:949-1001) batched over frames:
    p = calcActualPos() + tangents * phi -> closest point on the posed mesh -> faceIdx_ / vertexWeights_, against the
    oracle's IkTask (calc_actual_pos, calc_tangents) + float64 mesh projection on the oracle's own vertices."""
    from oracle import smpl_oracle as so
    from smplpp_b200 import api, synth
    faces0 = (np.asarray(params.face_indices, dtype=np.int64) - 1)
    _, face_idx, vw = marker_tasks
    tasks = api.IkTaskSet(smpl_gpu, face_idx, vposer=api.VPoserDecoder(vposer_params))
    B, n = 3, len(face_idx)
    beta, theta = synth.make_forward_inputs(B, 61)
    beta_shared = beta[0]
    rng = np.random.default_rng(8)
    phi = (rng.uniform(-0.01, 0.01, (B, n, 2)).astype(np.float32)) if with_phi else None
    w = torch.as_tensor(np.repeat(vw[None], B, axis=0), device="cuda:0").contiguous()
    state = torch.as_tensor(theta.reshape(B, 75), device="cuda:0").contiguous()
    fidx = torch.as_tensor(np.repeat(face_idx[None].astype(np.int32), B, axis=0), device="cuda:0").contiguous()
    opt = api.ik_options(normal_offset=0.015)
    tasks.reproject(opt, state, torch.as_tensor(beta_shared, device="cuda:0"), w, fidx,
                    dphi=None if phi is None else torch.as_tensor(phi, device="cuda:0"))
    face, wn = fidx.cpu().numpy(), w.cpu().numpy()
    agree = 0
    with torch.no_grad():
        r = so.smpl_launch(oracle_model, torch.as_tensor(np.repeat(beta_shared[None], B, 0)), torch.as_tensor(theta))
        for b in range(B):
            verts_o = r.vertices[b]
            pts = []
            for m in range(n):
                t = so.IkTask(int(face_idx[m]), normal_offset=0.015, vertex_weights=torch.as_tensor(vw[m]))
                p = t.calc_actual_pos(oracle_model, verts_o)
                if with_phi:
                    t.calc_tangents(oracle_model, verts_o)
                    p = p + torch.matmul(t.tangents, torch.as_tensor(phi[b, m]))
                pts.append(p.numpy())
            o_face, o_closest, o_sq, o_w = so.project_points_on_mesh(verts_o.numpy(), faces0, np.stack(pts))
            v64 = verts_o.numpy().astype(np.float64)
            for m in range(n):
                if face[b, m] == o_face[m]:
                    agree += 1
                    # weights are 1 / edge-length conditioned: compare the point they reproduce
                    got = (wn[b, m][:, None] * v64[faces0[face[b, m]]]).sum(0)
                    assert np.abs(got - o_closest[m]).max() < 1e-5
    assert agree >= 0.95 * B * n
    # on this mesh the 15 mm offset point often lands on another face: the re-seating must actually move attachments
    assert (face != face_idx[None]).any()
