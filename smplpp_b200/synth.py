"""Deterministic synthetic SMPL / VPoser parameters and mocap, with the official shapes.

The licensed model files (smpl_male.json, vposer_parameters.json) cannot be shipped, so every test and
benchmark runs on parameters generated here from fixed seeds.  The arrays carry exactly the
keys, shapes and conventions that the reference loads (src/SMPL.cpp:572-612, src/VPoser.cpp:185-237,
scripts/preprocess.py:88-121):

    face_indices        (13776, 3) int32, 1-BASED vertex ids
    shape_blend_shapes  (6890, 3, 10) float32
    pose_blend_shapes   (6890, 3, 207) float32
    vertices_template   (6890, 3) float32
    joint_regressor     (24, 6890) float32
    kinematic_tree      (2, 24) int64, row 0 = parents (root stored as 4294967295), row 1 = 0..23
    weights             (6890, 24) float32

This module is host-side numpy only; it never touches oracle/ or the CUDA extension.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass

import numpy as np

VERTEX_NUM = 6890
FACE_NUM = 13776
JOINT_NUM = 24
SHAPE_DIM = 10
POSE_DIM = 207
LATENT_DIM = 32
VPOSER_HIDDEN = 512
VPOSER_JOINTS = 21

# parents of the 24 SMPL joints (kinematic_tree[0]; src/toolbox/Tester.cpp:721-723)
PARENTS = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21], dtype=np.int64)

# T-pose joint table, metres, Y up, X to the body's left (SMPL-like proportions).
_JOINTS = np.array(
    [
        [0.000, 0.000, 0.000],  # 0 pelvis
        [0.070, -0.090, -0.005],  # 1 L hip
        [-0.070, -0.090, -0.005],  # 2 R hip
        [0.000, 0.110, -0.030],  # 3 spine1
        [0.105, -0.470, -0.010],  # 4 L knee
        [-0.105, -0.470, -0.010],  # 5 R knee
        [0.000, 0.245, -0.005],  # 6 spine2
        [0.090, -0.880, -0.045],  # 7 L ankle
        [-0.090, -0.880, -0.045],  # 8 R ankle
        [0.000, 0.300, 0.015],  # 9 spine3
        [0.115, -0.935, 0.075],  # 10 L foot
        [-0.115, -0.935, 0.075],  # 11 R foot
        [0.000, 0.510, -0.035],  # 12 neck
        [0.080, 0.420, -0.020],  # 13 L collar
        [-0.080, 0.420, -0.020],  # 14 R collar
        [0.005, 0.600, 0.020],  # 15 head
        [0.180, 0.450, -0.030],  # 16 L shoulder
        [-0.180, 0.450, -0.030],  # 17 R shoulder
        [0.440, 0.440, -0.045],  # 18 L elbow
        [-0.440, 0.440, -0.045],  # 19 R elbow
        [0.690, 0.445, -0.050],  # 20 L wrist
        [-0.690, 0.445, -0.050],  # 21 R wrist
        [0.775, 0.437, -0.062],  # 22 L hand
        [-0.775, 0.437, -0.062],  # 23 R hand
    ],
    dtype=np.float64,
)

# capsule radius around the bone parent(j) -> j
_BONE_RADIUS = np.array(
    [0.0, 0.115, 0.115, 0.135, 0.075, 0.075, 0.135, 0.055, 0.055, 0.140, 0.045, 0.045, 0.080, 0.090, 0.090,
     0.095, 0.070, 0.070, 0.048, 0.048, 0.038, 0.038, 0.034, 0.034]
)

# the OptiTrack Baseline-41 marker names of node/node.cpp:455-500 (std::map order = sorted) with an anchor
# (joint id, offset in metres) used to pick a face of the synthetic mesh.
_MARKER_ANCHORS = {
    "HeadTop": (15, (0.0, 0.16, 0.0)), "HeadFront": (15, (0.0, 0.08, 0.11)), "HeadSide": (15, (-0.09, 0.07, 0.0)),
    "Chest": (9, (0.0, 0.05, 0.14)), "WaistLFront": (0, (0.10, 0.03, 0.09)), "WaistRFront": (0, (-0.10, 0.03, 0.09)),
    "WaistLBack": (0, (0.08, 0.03, -0.11)), "WaistRBack": (0, (-0.08, 0.03, -0.11)),
    "BackTop": (12, (0.0, -0.04, -0.10)), "BackRight": (6, (-0.08, 0.0, -0.13)), "BackLeft": (6, (0.08, 0.0, -0.13)),
    "LShoulderTop": (16, (0.0, 0.07, 0.0)), "LShoulderBack": (16, (-0.03, 0.0, -0.08)),
    "LUArmHigh": (16, (0.12, 0.04, 0.02)), "LElbowOut": (18, (0.0, 0.02, -0.05)), "LWristIn": (20, (0.0, 0.0, 0.04)),
    "LWristOut": (20, (0.0, 0.0, -0.04)), "LHandOut": (22, (0.02, 0.03, 0.0)),
    "RShoulderTop": (17, (0.0, 0.07, 0.0)), "RShoulderBack": (17, (0.03, 0.0, -0.08)),
    "RUArmHigh": (17, (-0.12, 0.04, 0.02)), "RElbowOut": (19, (0.0, 0.02, -0.05)), "RWristIn": (21, (0.0, 0.0, 0.04)),
    "RWristOut": (21, (0.0, 0.0, -0.04)), "RHandOut": (23, (-0.02, 0.03, 0.0)),
    "LThigh": (1, (0.04, -0.20, 0.07)), "LKneeOut": (4, (0.06, 0.0, 0.0)), "LShin": (4, (0.0, -0.20, 0.05)),
    "LAnkleOut": (7, (0.05, 0.0, 0.0)), "LToeIn": (10, (-0.04, 0.0, 0.04)), "LToeOut": (10, (0.05, 0.0, 0.03)),
    "LToeTip": (10, (0.0, 0.0, 0.07)), "LHeel": (7, (0.0, -0.03, -0.06)),
    "RThigh": (2, (-0.04, -0.20, 0.07)), "RKneeOut": (5, (-0.06, 0.0, 0.0)), "RShin": (5, (0.0, -0.20, 0.05)),
    "RAnkleOut": (8, (-0.05, 0.0, 0.0)), "RToeIn": (11, (0.04, 0.0, 0.04)), "RToeOut": (11, (-0.05, 0.0, 0.03)),
    "RToeTip": (11, (0.0, 0.0, 0.07)), "RHeel": (8, (0.0, -0.03, -0.06)),
}
MARKER_NAMES = sorted(_MARKER_ANCHORS.keys())  # iteration order of the node's std::map (node/node.cpp:47)


@dataclass
class SmplParams:
    face_indices: np.ndarray
    shape_blend_shapes: np.ndarray
    pose_blend_shapes: np.ndarray
    vertices_template: np.ndarray
    joint_regressor: np.ndarray
    kinematic_tree: np.ndarray
    weights: np.ndarray

    def to_json(self, path: str) -> None:
        """Write the reference's model JSON (keys of src/SMPL.cpp:573-611).  ~55 MB of text."""

        def arr(a, fmt):
            a = np.asarray(a)
            if a.ndim == 1:
                return "[" + ",".join(fmt % x for x in a.tolist()) + "]"
            return "[" + ",".join(arr(s, fmt) for s in a) + "]"

        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "w") as f:
            # nested (V,3,D) layout is required by xt::from_json (shape(2) is checked, SMPL.cpp:579,586)
            f.write("{")
            f.write('"face_indices":' + arr(self.face_indices, "%d"))
            f.write(',"kinematic_tree":' + arr(self.kinematic_tree, "%d"))
            f.write(',"vertices_template":' + arr(self.vertices_template, "%.9g"))
            f.write(',"joint_regressor":' + arr(self.joint_regressor, "%.9g"))
            f.write(',"weights":' + arr(self.weights, "%.9g"))
            f.write(',"shape_blend_shapes":' + _nested3(self.shape_blend_shapes))
            f.write(',"pose_blend_shapes":' + _nested3(self.pose_blend_shapes))
            f.write("}")
        os.replace(tmp, path)


def _nested3(a: np.ndarray) -> str:
    v, k, d = a.shape
    rows = [",".join("%.9g" % x for x in r) for r in a.reshape(v * k, d).tolist()]
    out = []
    for i in range(v):
        out.append("[[" + "],[".join(rows[i * k:(i + 1) * k]) + "]]")
    return "[" + ",".join(out) + "]"


def _ray_capsule_far(o, d, a, b, r):
    """Largest t with |o + t d - segment(a,b)| <= r for unit directions d (N,3); -inf when missed."""
    n = d.shape[0]
    best = np.full(n, -np.inf)
    ab = b - a
    L2 = float(ab @ ab)

    def sphere(c):
        oc = o - c
        bq = d @ oc
        cq = float(oc @ oc) - r * r
        disc = bq * bq - cq
        t = np.where(disc >= 0, -bq + np.sqrt(np.maximum(disc, 0)), -np.inf)
        return t

    best = np.maximum(best, sphere(a))
    best = np.maximum(best, sphere(b))
    if L2 > 1e-12:
        u = ab / np.sqrt(L2)
        dperp = d - np.outer(d @ u, u)
        operp = (o - a) - ((o - a) @ u) * u
        A = np.einsum("ij,ij->i", dperp, dperp)
        B = dperp @ operp
        C = float(operp @ operp) - r * r
        disc = B * B - A * C
        ok = (disc >= 0) & (A > 1e-12)
        t = np.where(ok, (-B + np.sqrt(np.maximum(disc, 0))) / np.where(A > 1e-12, A, 1.0), -np.inf)
        s = ((o - a) @ u) + t * (d @ u)
        t = np.where(ok & (s >= 0) & (s <= np.sqrt(L2)), t, -np.inf)
        best = np.maximum(best, t)
    return best


def _radial_surface(dirs, center):
    far = np.full(dirs.shape[0], -np.inf)
    for j in range(1, JOINT_NUM):
        far = np.maximum(far, _ray_capsule_far(center, dirs, _JOINTS[PARENTS[j]], _JOINTS[j], _BONE_RADIUS[j]))
    # head ball and hands/feet tips
    far = np.maximum(far, _ray_capsule_far(center, dirs, _JOINTS[15], _JOINTS[15] + np.array([0, 0.09, 0.01]), 0.095))
    return far


def _point_segment_dist(p, a, b):
    ab = b - a
    t = np.clip(((p - a) @ ab) / max(float(ab @ ab), 1e-12), 0.0, 1.0)
    q = a + t[:, None] * ab
    return np.linalg.norm(p - q, axis=1)


_SMPL_CACHE = {}


def make_smpl_params(seed: int = 0) -> SmplParams:
    """Synthetic 'smpl_male-shaped' parameters.  Genus-0 surface => exactly 2V-4 = 13776 faces."""
    if seed in _SMPL_CACHE:
        return _SMPL_CACHE[seed]
    from scipy.spatial import ConvexHull

    rng = np.random.default_rng(seed)
    center = np.array([0.0, 0.20, -0.01])

    # directions: importance-sample so that thin limbs get a fair share of the 6890 vertices
    cand = rng.normal(size=(400000, 3))
    cand /= np.linalg.norm(cand, axis=1, keepdims=True)
    rad = _radial_surface(cand, center)
    wgt = np.clip(rad, 0.05, None) ** 2
    keep = rng.random(cand.shape[0]) < wgt / wgt.max()
    cand, rad = cand[keep], rad[keep]
    # thin out near-coincident surface points (keeps the smallest triangles well above fp32 noise)
    from scipy.spatial import cKDTree
    pts = center + rad[:, None] * cand
    drop = np.zeros(cand.shape[0], dtype=bool)
    for i, j in sorted(cKDTree(pts).query_pairs(0.006)):
        if not drop[i]:
            drop[j] = True
    dirs = cand[~drop][:VERTEX_NUM]
    assert dirs.shape[0] == VERTEX_NUM, dirs.shape
    # SMPL-like spatially coherent vertex numbering (the artist mesh numbers vertices region by region):
    # order by nearest bone, then by the position along that bone
    pts0 = center + _radial_surface(dirs, center)[:, None] * dirs
    dbone = np.stack([_point_segment_dist(pts0, _JOINTS[PARENTS[j]], _JOINTS[j]) if j else
                      np.linalg.norm(pts0 - _JOINTS[0], axis=1) for j in range(JOINT_NUM)], axis=1)
    near = dbone.argmin(axis=1)
    along = np.einsum("ij,ij->i", pts0 - _JOINTS[near], _JOINTS[near] - _JOINTS[np.maximum(PARENTS[near], 0)])
    dirs = dirs[np.lexsort((np.round(along, 2), near))]
    hull = ConvexHull(dirs)
    assert hull.vertices.shape[0] == VERTEX_NUM
    faces = hull.simplices.astype(np.int64)
    assert faces.shape == (FACE_NUM, 3), faces.shape
    # outward orientation
    p0, p1, p2 = dirs[faces[:, 0]], dirs[faces[:, 1]], dirs[faces[:, 2]]
    flip = np.einsum("ij,ij->i", np.cross(p1 - p0, p2 - p0), p0 + p1 + p2) < 0
    faces[flip] = faces[flip][:, [0, 2, 1]]
    order = np.lexsort((faces[:, 2], faces[:, 1], faces[:, 0]))
    faces = faces[order]

    verts = center + _radial_surface(dirs, center)[:, None] * dirs

    # skinning weights: softmax(-dist^2/sigma^2) over the <=4 nearest bones, exact zeros elsewhere
    dist = np.empty((VERTEX_NUM, JOINT_NUM))
    for j in range(JOINT_NUM):
        child = [c for c in range(JOINT_NUM) if PARENTS[c] == j]
        # bone "j" = segment from joint j to the mean of its children (or a short stub for leaves)
        end = _JOINTS[child].mean(axis=0) if child else _JOINTS[j] + 0.6 * (_JOINTS[j] - _JOINTS[PARENTS[j]])
        dist[:, j] = _point_segment_dist(verts, _JOINTS[j], end)
    sigma = 0.06
    logit = -(dist / sigma) ** 2
    idx = np.argsort(dist, axis=1)[:, :4]
    weights = np.zeros((VERTEX_NUM, JOINT_NUM))
    rows = np.arange(VERTEX_NUM)[:, None]
    sel = logit[rows, idx]
    sel = np.exp(sel - sel.max(axis=1, keepdims=True))
    sel[sel < 1e-4 * sel.max(axis=1, keepdims=True)] = 0.0
    weights[rows, idx] = sel / sel.sum(axis=1, keepdims=True)

    # joint regressor: non-negative rows summing to 1, supported near each joint, and reproducing the
    # joint table (least-norm correction inside the support keeps rows sparse and mostly non-negative)
    jreg = np.zeros((JOINT_NUM, VERTEX_NUM))
    for j in range(JOINT_NUM):
        dj = np.linalg.norm(verts - _JOINTS[j], axis=1)
        radius = 0.08
        sup = np.where(dj < radius)[0]
        while sup.shape[0] < 24:
            radius *= 1.25
            sup = np.where(dj < radius)[0]
        w = np.exp(-(dj[sup] / (0.6 * radius)) ** 2)
        w /= w.sum()
        M = np.concatenate([verts[sup].T, np.ones((1, sup.shape[0]))], axis=0)  # (4, n)
        resid = np.concatenate([_JOINTS[j] - verts[sup].T @ w, [0.0]])
        w = w + M.T @ np.linalg.solve(M @ M.T + 1e-12 * np.eye(4), resid)
        jreg[j, sup] = w
    shape_basis = np.empty((VERTEX_NUM, 3, SHAPE_DIM))
    # smooth, low-frequency shape directions (cm scale like the real model)
    freq = rng.normal(size=(SHAPE_DIM, 3, 3)) * 2.5
    phase = rng.uniform(0, 2 * np.pi, size=(SHAPE_DIM, 3))
    amp = 0.012 / (1.0 + 0.25 * np.arange(SHAPE_DIM))
    for i in range(SHAPE_DIM):
        for k in range(3):
            shape_basis[:, k, i] = amp[i] * np.sin(verts @ freq[i, k] + phase[i, k])
        # component 0/1 also scale the body (height / girth), like the leading real shape components
    shape_basis[:, :, 0] += 0.03 * verts * np.array([0.3, 1.0, 0.3])
    shape_basis[:, :, 1] += 0.04 * (verts - center) * np.array([1.0, 0.0, 1.0])
    pose_basis = rng.normal(scale=0.002, size=(VERTEX_NUM, 3, POSE_DIM))

    tree = np.stack([PARENTS.copy(), np.arange(JOINT_NUM, dtype=np.int64)])
    tree[0, 0] = 4294967295
    params = SmplParams(
        face_indices=(faces + 1).astype(np.int32),
        shape_blend_shapes=shape_basis.astype(np.float32),
        pose_blend_shapes=pose_basis.astype(np.float32),
        vertices_template=verts.astype(np.float32),
        joint_regressor=jreg.astype(np.float32),
        kinematic_tree=tree,
        weights=weights.astype(np.float32),
    )
    _SMPL_CACHE[seed] = params
    return params


def make_vposer_params(seed: int = 1) -> dict:
    """decoder_net.{0,3,5}.{weight,bias} with the shapes checked by src/VPoser.cpp:185-237."""
    rng = np.random.default_rng(seed)
    H = VPOSER_HIDDEN
    out = 6 * VPOSER_JOINTS

    def lin(o, i):
        return rng.normal(scale=1.0 / np.sqrt(i), size=(o, i)).astype(np.float32)

    return {
        "decoder_net.0.weight": lin(H, LATENT_DIM),
        "decoder_net.0.bias": rng.normal(scale=0.1, size=H).astype(np.float32),
        "decoder_net.3.weight": lin(H, H),
        "decoder_net.3.bias": rng.normal(scale=0.1, size=H).astype(np.float32),
        "decoder_net.5.weight": lin(out, H),
        "decoder_net.5.bias": rng.normal(scale=0.1, size=out).astype(np.float32),
    }


def vposer_to_json(params: dict, path: str) -> None:
    with open(path, "w") as f:
        json.dump({k: np.asarray(v, dtype=np.float64).tolist() for k, v in params.items()}, f)


def make_forward_inputs(batch: int, seed: int):
    """beta (B,10) ~ N(0,1); theta (B,25,3): row 0 translation U(-1,1), row 1 root N(0,0.5^2), rows 2..24
    N(0,0.3^2) rad."""
    rng = np.random.default_rng(seed)
    beta = rng.normal(size=(batch, SHAPE_DIM)).astype(np.float32)
    theta = np.empty((batch, JOINT_NUM + 1, 3), dtype=np.float32)
    theta[:, 0] = rng.uniform(-1, 1, size=(batch, 3))
    theta[:, 1] = rng.normal(scale=0.5, size=(batch, 3))
    theta[:, 2:] = rng.normal(scale=0.3, size=(batch, JOINT_NUM - 1, 3))
    return beta, theta


def make_marker_tasks(params: SmplParams, seed: int = 2):
    """41 marker attachments: (names sorted, faceIdx (41,) 0-based row of face_indices, vertexWeights (41,3))."""
    rng = np.random.default_rng(seed)
    faces = params.face_indices.astype(np.int64) - 1
    verts = params.vertices_template.astype(np.float64)
    cent = verts[faces].mean(axis=1)
    tri = verts[faces]
    area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    edge = np.max([np.linalg.norm(tri[:, a] - tri[:, b], axis=1) for a, b in ((0, 1), (1, 2), (2, 0))], axis=0)
    good = (area > np.quantile(area, 0.3)) & (edge < 0.05)  # well-shaped faces only
    face_idx = np.empty(len(MARKER_NAMES), dtype=np.int64)
    used = set()
    for m, name in enumerate(MARKER_NAMES):
        j, off = _MARKER_ANCHORS[name]
        target = _JOINTS[j] + np.asarray(off)
        orderf = np.argsort(np.linalg.norm(cent - target, axis=1))
        for f in orderf:
            if int(f) not in used and good[f]:
                used.add(int(f))
                face_idx[m] = f
                break
    w = rng.dirichlet(np.ones(3), size=len(MARKER_NAMES)).astype(np.float32)
    w = np.clip(w, 0.05, None)
    w = (w / w.sum(axis=1, keepdims=True)).astype(np.float32)
    return list(MARKER_NAMES), face_idx, w


def make_motion(frames: int, seed: int, z_up: bool = True):
    """Ground-truth theta(t) (frames,25,3) for a 'sample_walk-shaped' clip: smooth random walk at 120 Hz.

    Root orientation starts at rpy=(1.57,0,3.14)-like axis-angle when z_up (launch/smplpp.launch:41), the
    translation drifts inside the capture volume of data/sample_walk.c3d."""
    rng = np.random.default_rng(seed)
    t = np.arange(frames) / 120.0
    theta = np.zeros((frames, JOINT_NUM + 1, 3))
    # low-frequency sinusoid mixtures per joint dof
    nh = 3
    for j in range(1, JOINT_NUM + 1):
        amp = 0.12 if j == 1 else 0.25
        if j in (23, 24):  # hands: small
            amp = 0.05
        a = rng.normal(scale=amp / nh, size=(nh, 3))
        f = rng.uniform(0.2, 1.6, size=(nh, 3))
        p = rng.uniform(0, 2 * np.pi, size=(nh, 3))
        theta[:, j] = np.sum(a[None] * np.sin(2 * np.pi * f[None] * t[:, None, None] + p[None]), axis=1)
    if z_up:
        # rotation taking SMPL's Y-up rest pose to a Z-up world, facing -Y: axis-angle of Rx(pi/2)
        theta[:, 1] += np.array([np.pi / 2, 0.0, 0.0])
        theta[:, 0] = np.array([-0.7, 0.3, 0.95])
        theta[:, 0, 0] += 0.9 * np.sin(2 * np.pi * 0.11 * t)
        theta[:, 0, 1] += 1.2 * np.sin(2 * np.pi * 0.07 * t + 0.4)
        theta[:, 0, 2] += 0.03 * np.sin(2 * np.pi * 1.8 * t)
    else:
        theta[:, 0] = 0.3 * np.stack([np.sin(0.5 * t), 0.1 * np.cos(0.7 * t), np.sin(0.3 * t + 1)], axis=1)
    return theta.astype(np.float32)


def make_marker_noise(frames: int, markers: int, seed: int, sigma: float = 1e-3, dropout: float = 0.034):
    """(noise (frames,markers,3) float32 ~ N(0, sigma^2), valid (frames,markers) bool with 3.4 % dropout)."""
    rng = np.random.default_rng(seed)
    noise = rng.normal(scale=sigma, size=(frames, markers, 3)).astype(np.float32)
    valid = rng.random((frames, markers)) >= dropout
    return noise, valid


def initial_theta(z_up: bool = True) -> np.ndarray:
    """Common initial pose for batched IK (the reference warm-starts serially; batched frames cannot)."""
    th = np.zeros((JOINT_NUM + 1, 3), dtype=np.float32)
    if z_up:
        th[1] = [np.pi / 2, 0.0, 0.0]
        th[0] = [-0.7, 0.3, 0.95]
    return th
