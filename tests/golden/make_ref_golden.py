"""Generates golden input/output vectors by running the UNMODIFIED reference sources (oracle/_ref, built by
`make -C oracle`) on seeded synthetic inputs.  Run in the build container; the .npz files are committed and
are what the CPU and GPU test-suites compare against when oracle/_ref is not at hand.

    python tests/golden/make_ref_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import ref_lib  # noqa: E402
from smplpp_b200 import synth  # noqa: E402


def rotations_for_property_test(seed=3):
    """The input families of tests/src/TestVPoser.cpp:45-70: identity, +-{0,1e-12..1e-1} around {0,pi/4,pi/2,pi}
    about each axis, and random rotations."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    vecs = [np.zeros(3)]
    for base in (0.0, np.pi / 4, np.pi / 2, np.pi):
        for eps in (0.0, 1e-12, 1e-10, 1e-8, 1e-6, 1e-4, 1e-3, 1e-2, 1e-1):
            for sgn in (1.0, -1.0):
                for ax in np.eye(3):
                    vecs.append((base + sgn * eps) * ax)
                    vecs.append(-(base + sgn * eps) * ax)
    for _ in range(300):
        a = rng.normal(size=3)
        a /= np.linalg.norm(a)
        vecs.append(a * rng.uniform(0, np.pi))
    for _ in range(60):  # close to pi about random axes
        a = rng.normal(size=3)
        a /= np.linalg.norm(a)
        vecs.append(a * (np.pi - rng.choice([0, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2])))
    vecs = np.asarray(vecs)
    return vecs.astype(np.float32), Rotation.from_rotvec(vecs).as_matrix().astype(np.float32)


def main():
    ref_lib.set_num_threads(1)  # deterministic summation order
    params = synth.make_smpl_params(0)
    ref = ref_lib.RefSMPL(ref_lib.model_json_path(0))

    # --- forward (config 1 inputs: rng(10)) ---
    beta, theta = synth.make_forward_inputs(2, 10)
    out = ref.forward(beta, theta)
    names, face_idx, vw = synth.make_marker_tasks(params)
    faces0 = params.face_indices.astype(np.int64) - 1
    vert_idx = np.unique(faces0[face_idx].reshape(-1))
    ref.forward(beta[:1], theta[:1], want=("vertices",))
    fn, vn = ref.normals(face_idx, vert_idx)
    np.savez_compressed(os.path.join(OUT, "ref_forward.npz"), beta=beta, theta=theta, vertices=out["vertices"],
                        joints=out["joints"], rest_shape=out["rest_shape"], normal_face_idx=face_idx,
                        normal_vert_idx=vert_idx, face_normals=fn, vertex_normals=vn)

    # --- VPoser ---
    vp = ref_lib.RefVPoser(ref_lib.vposer_json_path(1))
    rng = np.random.default_rng(30)
    latent = (rng.normal(size=(16, 32)) * np.linspace(0.2, 3.0, 16)[:, None]).astype(np.float32)
    aa = vp.forward(latent)
    _, jac = vp.forward(latent[:3], jacobian=True)
    aa_in, rot_in = rotations_for_property_test()
    aa_out, aa_grad = ref_lib.rotmat_to_axis_angle(rot_in, grad=True)
    np.savez_compressed(os.path.join(OUT, "ref_vposer.npz"), latent=latent, axis_angle=aa, jacobian=jac,
                        prop_rotvec=aa_in, prop_rotmat=rot_in, prop_axis_angle=aa_out, prop_grad=aa_grad)

    # --- IK iterations (one frame each) ---
    gt = synth.make_motion(8, 20)
    b0 = (np.random.default_rng(5).normal(size=10) * 0.5).astype(np.float32)
    n = len(names)
    # targets: marker positions of the ground-truth pose (reference forward + IkTask::calcActualPos) + 1 mm noise
    noise, valid = synth.make_marker_noise(8, n, 21)
    x_gt = gt[3].reshape(-1)
    r0 = ref.ik_iteration(x_gt, b0, face_idx, vw, np.zeros((n, 3), np.float32), normal_task_weight=0.0, phi_limit=0.0,
                          normal_offset=0.015, update_state=False)
    # actual_pos is evaluated on the launched (= ground-truth) mesh; with phi_limit = 0 delta_phi = 0
    target = r0["actual_pos"] + noise[3]
    x0 = gt[0].reshape(-1).copy()  # start from a nearby frame of the clip
    modes = {
        "motion": dict(normal_task_weight=0.0, phi_limit=0.0, normal_offset=0.015, optimize_beta=False,
                       pos_task_weight=valid[3].astype(np.float64)),
        "body": dict(normal_task_weight=0.0, phi_limit=0.04, normal_offset=0.015, optimize_beta=True),
        "interactive": dict(normal_task_weight=1.0, phi_limit=0.0, normal_offset=0.0, optimize_beta=False),
        "llt": dict(normal_task_weight=0.0, phi_limit=0.0, normal_offset=0.0, optimize_beta=False, enable_qp=False),
    }
    ik = dict(face_idx=face_idx, vertex_weights_in=vw, target_pos=target, theta_in=x0, beta_in=b0)
    for mode, kw in modes.items():
        tp = target.copy()
        if "pos_task_weight" in kw:
            tp[kw["pos_task_weight"] == 0] = 0.0  # node.cpp:682-683
        r = ref.ik_iteration(x0, b0, face_idx, vw, tp, **kw)
        ik[mode + "_target"] = tp
        if "pos_task_weight" in kw:
            ik[mode + "_pos_task_weight"] = kw["pos_task_weight"]
        for k in ("e", "b", "delta"):
            ik["%s_%s" % (mode, k)] = r[k]
        ik[mode + "_J"] = r["J"].astype(np.float32)  # fp32 gradients cast to fp64 in the reference
        ik[mode + "_theta_out"] = r["theta_state"]
        ik[mode + "_beta_out"] = r["beta"]
        ik[mode + "_vertex_weights_out"] = r["vertex_weights"]
        ik[mode + "_actual_pos"] = r["actual_pos"]
        print(mode, "|e|", np.linalg.norm(r["e"]), "|delta|", np.linalg.norm(r["delta"]))
    # VPoser mode (44-dim state)
    xv = np.zeros(44, np.float32)
    xv[0:3] = x0[0:3]
    xv[3:6] = x0[3:6]
    xv[6:38] = np.random.default_rng(31).normal(size=32).astype(np.float32) * 0.5
    xv[38:44] = 0.01
    r = ref.ik_iteration(xv, b0, face_idx, vw, target, normal_task_weight=0.0, phi_limit=0.0, normal_offset=0.015,
                         vposer=vp)
    ik["vposer_theta_in"] = xv
    for k in ("e", "b", "delta"):
        ik["vposer_%s" % k] = r[k]
    ik["vposer_J"] = r["J"].astype(np.float32)
    ik["vposer_theta_out"] = r["theta_state"]
    ik["vposer_vertex_weights_out"] = r["vertex_weights"]
    print("vposer |e|", np.linalg.norm(r["e"]), "|delta|", np.linalg.norm(r["delta"]))
    np.savez_compressed(os.path.join(OUT, "ref_ik.npz"), **ik)
    for f in ("ref_forward.npz", "ref_vposer.npz", "ref_ik.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
