"""Parity of the header-only C++ facade (include/smplpp_b200/smplpp.hpp) through its own binary: tests/cpp/facade_parity
loads tensors of the COMPILED REFERENCE (tests/golden/*.npz, dumped here as raw files) and drives smplpp::SMPL, the four
module classes, smplpp::IkTask, smplpp::IkTaskSet::step + getJacobian and smplpp::VPoserDecoder::forward (the C++ object
API of include/smplpp/{SMPL,IkTask,VPoser,BlendShape,...}.h) with the north_star tolerances."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "facade_parity")
SMOKE = os.path.join(ROOT, "tests", "cpp", "facade_smoke")


def _dump(d, name, a, dtype):
    np.ascontiguousarray(a, dtype=dtype).tofile(os.path.join(d, name))


@pytest.mark.gpu
def test_cpp_facade_parity_vs_reference_goldens(tmp_path, params, vposer_params, golden_forward, golden_ik, golden_vposer):
    from smplpp_b200 import synth
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(BIN)])
    d = str(tmp_path)
    np.savez(os.path.join(d, "model.npz"), face_indices=params.face_indices, shape_blend_shapes=params.shape_blend_shapes,
             pose_blend_shapes=params.pose_blend_shapes, vertices_template=params.vertices_template,
             joint_regressor=params.joint_regressor, kinematic_tree=params.kinematic_tree, weights=params.weights)
    synth.vposer_to_json(vposer_params, os.path.join(d, "vposer.json"))
    for key in ("shape_blend_shapes", "pose_blend_shapes", "vertices_template", "joint_regressor", "weights"):
        _dump(d, key + ".f32", getattr(params, key), np.float32)
    _dump(d, "kinematic_tree.i64", params.kinematic_tree, np.int64)
    g = golden_forward
    for key in ("beta", "theta", "vertices", "joints", "rest_shape"):
        _dump(d, "fwd_%s.f32" % key, g[key], np.float32)
    _dump(d, "normal_face_idx.i64", g["normal_face_idx"], np.int64)
    _dump(d, "normal_vert_idx.i64", g["normal_vert_idx"], np.int64)
    _dump(d, "face_normals.f32", g["face_normals"], np.float32)
    _dump(d, "vertex_normals.f32", g["vertex_normals"], np.float32)
    k = golden_ik
    _dump(d, "ik_face_idx.i64", k["face_idx"], np.int64)
    for key in ("theta_in", "beta_in", "vertex_weights_in", "motion_actual_pos", "motion_vertex_weights_out", "motion_target",
                "motion_pos_task_weight", "motion_J", "motion_theta_out"):
        _dump(d, "ik_%s.f32" % key, k[key], np.float32)
    _dump(d, "ik_motion_e.f64", k["motion_e"], np.float64)
    _dump(d, "vposer_latent.f32", golden_vposer["latent"], np.float32)
    _dump(d, "vposer_axis_angle.f32", golden_vposer["axis_angle"], np.float32)
    r = subprocess.run([BIN, d], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "facade parity: OK" in r.stdout and "FAIL" not in r.stdout


@pytest.mark.gpu
def test_cpp_facade_smoke():
    if not os.path.exists(SMOKE):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(SMOKE)])
    r = subprocess.run([SMOKE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_cpp_facade_builds_without_cuda_headers():
    """The facade is plain C++17 over the C ABI: g++ alone compiles and links both binaries (done by build())."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])
    assert os.path.exists(BIN) and os.path.exists(SMOKE)
