"""TEST INFRASTRUCTURE — ctypes binding of oracle/_ref/libsmplpp_ref.so (the UNMODIFIED reference sources
compiled by oracle/Makefile + the C harness oracle/ref_harness.cpp).

Used only by tests/, tests/golden/make_ref_golden.py, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing here reads /root/reference at run time: the shared object is prebuilt and
travels with the repo snapshot; the model JSON it loads is generated from smplpp_b200.synth.
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libsmplpp_ref.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def available() -> bool:
    return os.path.exists(LIB_PATH)


class _IkOptions(C.Structure):
    _fields_ = [("enableVposer", C.c_int32), ("optimizeBeta", C.c_int32), ("enableQp", C.c_int32),
                ("nTasks", C.c_int32), ("updateState", C.c_int32), ("reserved", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        import torch  # noqa: F401  (loads libtorch/libc10 so the rpath-less case also resolves)
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_last_error.restype = C.c_char_p
        _lib.ref_get_num_threads.restype = C.c_int
    return _lib


def _check(rc, ok=(0,)):
    if rc not in ok:
        raise RuntimeError("reference harness failed (%d): %s" % (rc, lib().ref_last_error().decode()))
    return rc


def _opt(a, dtype):
    if a is None:
        return None
    assert a.dtype == dtype and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def set_num_threads(n: int):
    lib().ref_set_num_threads(int(n))


def get_num_threads() -> int:
    return int(lib().ref_get_num_threads())


def model_json_path(seed: int = 0) -> str:
    """Synthetic model JSON for SMPL::init, cached under the system temp dir (55 MB of text)."""
    from smplpp_b200 import synth
    path = os.path.join(tempfile.gettempdir(), "smplpp_b200_synth_smpl_seed%d.json" % seed)
    if not os.path.exists(path):
        synth.make_smpl_params(seed).to_json(path)
    return path


def vposer_json_path(seed: int = 1) -> str:
    from smplpp_b200 import synth
    path = os.path.join(tempfile.gettempdir(), "smplpp_b200_synth_vposer_seed%d.json" % seed)
    if not os.path.exists(path):
        synth.vposer_to_json(synth.make_vposer_params(seed), path)
    return path


class RefSMPL:
    """smplpp::SMPL (src/SMPL.cpp) on libtorch CPU."""

    def __init__(self, json_path: str):
        self._h = C.c_void_p()
        _check(lib().ref_smpl_create(json_path.encode(), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ref_smpl_destroy(self._h)
            self._h = None

    def forward(self, beta, theta, batched=False, no_grad=True, want=("vertices", "joints", "rest_shape")):
        beta = np.ascontiguousarray(beta, dtype=np.float32)
        theta = np.ascontiguousarray(theta, dtype=np.float32)
        b = beta.shape[0]
        out = {
            "vertices": np.empty((b, 6890, 3), np.float32) if "vertices" in want else None,
            "joints": np.empty((b, 24, 3), np.float32) if "joints" in want else None,
            "rest_shape": np.empty((b, 6890, 3), np.float32) if "rest_shape" in want else None,
        }
        _check(lib().ref_smpl_forward(self._h, C.c_int64(b), int(batched), int(no_grad), _opt(beta, np.float32),
                                      _opt(theta, np.float32), _opt(out["vertices"], np.float32),
                                      _opt(out["joints"], np.float32), _opt(out["rest_shape"], np.float32)))
        return out

    def normals(self, face_idx, vert_idx):
        face_idx = np.ascontiguousarray(face_idx, dtype=np.int64)
        vert_idx = np.ascontiguousarray(vert_idx, dtype=np.int64)
        fn = np.empty((face_idx.shape[0], 3), np.float32)
        vn = np.empty((vert_idx.shape[0], 3), np.float32)
        _check(lib().ref_smpl_normals(self._h, C.c_int64(face_idx.shape[0]), _opt(face_idx, np.int64),
                                      _opt(fn, np.float32), C.c_int64(vert_idx.shape[0]), _opt(vert_idx, np.int64),
                                      _opt(vn, np.float32)))
        return fn, vn

    def ik_iteration(self, theta_state, beta, face_idx, vertex_weights, target_pos, target_normal=None,
                     pos_task_weight=None, normal_task_weight=None, phi_limit=None, normal_offset=None,
                     vposer: "RefVPoser" = None, optimize_beta=False, enable_qp=True, update_state=True):
        """Restated node/node.cpp:705-968 on the compiled reference objects (oracle/ref_harness.cpp)."""
        n = len(face_idx)
        theta_state = np.ascontiguousarray(theta_state, dtype=np.float32).reshape(-1).copy()
        beta = np.ascontiguousarray(beta, dtype=np.float32).reshape(-1).copy()
        theta_dim = theta_state.shape[0]
        dim = theta_dim + 2 * n + (10 if optimize_beta else 0)
        face_idx = np.ascontiguousarray(face_idx, dtype=np.int64)
        vw = np.ascontiguousarray(vertex_weights, dtype=np.float32).copy()
        tp = np.ascontiguousarray(target_pos, dtype=np.float32)
        tn = np.ascontiguousarray(np.tile([0, 0, 1.0], (n, 1)) if target_normal is None else target_normal,
                                  dtype=np.float32)

        def vec(x, default):
            return np.ascontiguousarray(np.full(n, default) if x is None else np.broadcast_to(x, (n,)),
                                        dtype=np.float64)

        pw, nw = vec(pos_task_weight, 1.0), vec(normal_task_weight, 1.0)
        pl, no = vec(phi_limit, 0.04), vec(normal_offset, 0.0)
        e = np.empty(4 * n)
        J = np.empty((4 * n, dim))
        A = np.empty((dim, dim))
        bb = np.empty(dim)
        delta = np.empty(dim)
        actual = np.empty((n, 3), np.float32)
        opt = _IkOptions(int(vposer is not None), int(optimize_beta), int(enable_qp), n, int(update_state), 0)
        rc = _check(lib().ref_ik_iteration(
            self._h, vposer._h if vposer is not None else None, C.byref(opt), _opt(theta_state, np.float32),
            _opt(beta, np.float32), _opt(face_idx, np.int64), _opt(vw, np.float32), _opt(tp, np.float32),
            _opt(tn, np.float32), _opt(pw, np.float64), _opt(nw, np.float64), _opt(pl, np.float64),
            _opt(no, np.float64), _opt(e, np.float64), _opt(J, np.float64), _opt(A, np.float64),
            _opt(bb, np.float64), _opt(delta, np.float64), _opt(actual, np.float32)), ok=(0, 1))
        return dict(e=e, J=J, A=A, b=bb, delta=delta, theta_state=theta_state, beta=beta, vertex_weights=vw,
                    actual_pos=actual, skipped=(rc == 1))


class RefVPoser:
    """smplpp::VPoserDecoder (src/VPoser.cpp) on libtorch CPU."""

    def __init__(self, json_path: str):
        self._h = C.c_void_p()
        _check(lib().ref_vposer_create(json_path.encode(), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ref_vposer_destroy(self._h)
            self._h = None

    def forward(self, latent, jacobian=False):
        latent = np.ascontiguousarray(latent, dtype=np.float32)
        b = latent.shape[0]
        aa = np.empty((b, 21, 3), np.float32)
        jac = np.empty((b, 63, 32), np.float32) if jacobian else None
        _check(lib().ref_vposer_forward(self._h, C.c_int64(b), _opt(latent, np.float32), _opt(aa, np.float32),
                                        _opt(jac, np.float32)))
        return (aa, jac) if jacobian else aa


def rotmat_to_axis_angle(rot, grad=False):
    rot = np.ascontiguousarray(rot, dtype=np.float32)
    n = rot.shape[0]
    aa = np.empty((n, 3), np.float32)
    g = np.empty((n, 3, 3), np.float32) if grad else None
    _check(lib().ref_rotmat_to_axis_angle(C.c_int64(n), _opt(rot, np.float32), _opt(aa, np.float32),
                                          _opt(g, np.float32)))
    return (aa, g) if grad else aa


def triangle_vertex_weights(pos, tri):
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    tri = np.ascontiguousarray(tri, dtype=np.float32)
    w = np.empty(3, np.float32)
    _check(lib().ref_triangle_vertex_weights(_opt(pos, np.float32), _opt(tri, np.float32), _opt(w, np.float32)))
    return w


def blend_shape(beta, theta, shape_basis, pose_basis):
    beta = np.ascontiguousarray(beta, np.float32)
    theta = np.ascontiguousarray(theta, np.float32)
    sb = np.ascontiguousarray(shape_basis, np.float32)
    pb = np.ascontiguousarray(pose_basis, np.float32)
    b, v = beta.shape[0], sb.shape[0]
    s = np.empty((b, v, 3), np.float32)
    p = np.empty((b, v, 3), np.float32)
    r = np.empty((b, 24, 3, 3), np.float32)
    _check(lib().ref_blend_shape(C.c_int64(b), C.c_int64(v), _opt(beta, np.float32), _opt(theta, np.float32),
                                 _opt(sb, np.float32), _opt(pb, np.float32), _opt(s, np.float32),
                                 _opt(p, np.float32), _opt(r, np.float32)))
    return s, p, r


def joint_regression(templ, jreg, shape_bs, pose_bs):
    templ = np.ascontiguousarray(templ, np.float32)
    jreg = np.ascontiguousarray(jreg, np.float32)
    s = np.ascontiguousarray(shape_bs, np.float32)
    p = np.ascontiguousarray(pose_bs, np.float32)
    b, v = s.shape[0], templ.shape[0]
    rest = np.empty((b, v, 3), np.float32)
    joints = np.empty((b, 24, 3), np.float32)
    _check(lib().ref_joint_regression(C.c_int64(b), C.c_int64(v), _opt(templ, np.float32), _opt(jreg, np.float32),
                                      _opt(s, np.float32), _opt(p, np.float32), _opt(rest, np.float32),
                                      _opt(joints, np.float32)))
    return rest, joints


def world_transformation(kine_tree, joints, pose_rot):
    kt = np.ascontiguousarray(kine_tree, np.int64)
    j = np.ascontiguousarray(joints, np.float32)
    r = np.ascontiguousarray(pose_rot, np.float32)
    b = j.shape[0]
    out = np.empty((b, 24, 4, 4), np.float32)
    _check(lib().ref_world_transformation(C.c_int64(b), _opt(kt, np.int64), _opt(j, np.float32),
                                          _opt(r, np.float32), _opt(out, np.float32)))
    return out


def linear_blend_skinning(weights, rest, transforms, root_pos):
    w = np.ascontiguousarray(weights, np.float32)
    r = np.ascontiguousarray(rest, np.float32)
    t = np.ascontiguousarray(transforms, np.float32)
    rp = np.ascontiguousarray(root_pos, np.float32)
    b, v = r.shape[0], w.shape[0]
    out = np.empty((b, v, 3), np.float32)
    _check(lib().ref_linear_blend_skinning(C.c_int64(b), C.c_int64(v), _opt(w, np.float32), _opt(r, np.float32),
                                           _opt(t, np.float32), _opt(rp, np.float32), _opt(out, np.float32)))
    return out
