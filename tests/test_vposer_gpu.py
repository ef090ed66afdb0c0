"""GPU parity tests of the VPoser decoder kernels (MLP, 6D -> rotation, rotation -> axis-angle, 63x32 Jacobian)."""
import numpy as np
import pytest
import torch

from conftest import TOL_JACOBIAN_REL

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.fixture(scope="module")
def vposer_gpu(vposer_params):
    from smplpp_b200 import api
    return api.VPoserDecoder(vposer_params)


def test_decoder_vs_reference_golden(vposer_gpu, golden_vposer):
    g = golden_vposer
    aa, jac = vposer_gpu.forward(g["latent"], jacobian=True)
    aa, jac = aa.cpu().numpy(), jac.cpu().numpy()
    # B2 branch w * theta / (2 sin theta) amplifies fp32 rounding towards 172 deg: 2e-5 budget
    assert np.abs(aa - g["axis_angle"]).max() < 2e-5
    assert np.abs(vposer_gpu.forward(g["latent"]).cpu().numpy() - aa).max() == 0.0  # fwd-only kernel agrees
    for i in range(g["jacobian"].shape[0]):
        ref = g["jacobian"][i]
        assert np.abs(jac[i] - ref).max() / np.abs(ref).max() <= TOL_JACOBIAN_REL


def test_decoder_vs_oracle_batch(vposer_gpu, oracle_vposer):
    rng = np.random.default_rng(32)
    z = (rng.normal(size=(301, 32)) * rng.uniform(0.1, 4.0, size=(301, 1))).astype(f32)
    with torch.no_grad():
        ref = oracle_vposer.forward(torch.as_tensor(z)).numpy()
    aa = vposer_gpu.forward(z).cpu().numpy()
    # axis-angle is ill-conditioned towards pi (w theta / (2 sin theta)): compare the axis-angle
    # where the angle is below 2.6 rad and the ROTATIONS everywhere
    from scipy.spatial.transform import Rotation
    ang = np.linalg.norm(ref.reshape(-1, 3), axis=1)
    err = np.abs(aa - ref).reshape(-1, 3).max(axis=1)
    assert err[ang < 2.6].max() < 5e-5
    r_gpu = Rotation.from_rotvec(aa.reshape(-1, 3).astype(np.float64)).as_matrix()
    r_ref = Rotation.from_rotvec(ref.reshape(-1, 3).astype(np.float64)).as_matrix()
    derr = np.abs(r_gpu - r_ref).max(axis=(1, 2))
    assert derr[ang < 3.0].max() < 1e-4
    assert np.isfinite(aa).all() and derr.max() < 5e-3  # within 0.14 rad of pi: sqrt(s + eps) branch, VPoser.cpp:54-60
    zt = torch.as_tensor(z[:4]).requires_grad_(True)
    out = oracle_vposer.forward(zt).reshape(4, 63)
    _, jac = vposer_gpu.forward(z[:4], jacobian=True)
    jac = jac.cpu().numpy()
    ang4 = ang.reshape(-1, 21)[:4]
    for f in range(4):
        ref_j = np.stack([torch.autograd.grad(out[f, r], zt, retain_graph=True)[0][f].numpy() for r in range(63)])
        d = np.abs(jac[f] - ref_j).reshape(21, 3, 32).max(axis=(1, 2))
        scale = np.abs(ref_j).max()
        # the random synthetic decoder produces joints within 0.01 rad of pi, where d(axis-angle)/dR ~ 1 / sin(theta)
        # amplifies fp32 rounding of BOTH implementations (VPoser.cpp:37-60): the 1e-4 bound holds below 2.6 rad
        assert d[ang4[f] < 2.6].max() / scale <= TOL_JACOBIAN_REL
        assert d.max() / scale <= 5e-2


def test_rotmat_to_axis_angle_golden_and_property(golden_vposer):
    """Reference golden + the property test of tests/src/TestVPoser.cpp:16-70 on the CUDA kernel."""
    from scipy.spatial.transform import Rotation
    from smplpp_b200 import api
    g = golden_vposer
    got = api.convertRotMatToAxisAngle(g["prop_rotmat"]).cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got - g["prop_axis_angle"]).max() < 5e-6
    truth = Rotation.from_matrix(g["prop_rotmat"].astype(np.float64)).as_rotvec()
    err = np.minimum(np.linalg.norm(got - truth, axis=1), np.linalg.norm(got + truth, axis=1))
    assert (err[: err.shape[0] - 60] < 5e-3).all()


def test_decoder_jacobian_finite_difference(vposer_gpu):
    """Size-independent property: J matches central differences of the kernel's own forward."""
    rng = np.random.default_rng(33)
    z = rng.normal(size=(3, 32)).astype(f32)
    _, jac = vposer_gpu.forward(z, jacobian=True)
    jac = jac.cpu().numpy()
    h = 2e-3
    bad = total = 0
    for t in range(0, 32, 5):
        zp, zm = z.copy(), z.copy()
        zp[:, t] += h
        zm[:, t] -= h
        fd = (vposer_gpu.forward(zp).cpu().numpy() - vposer_gpu.forward(zm).cpu().numpy()).reshape(3, 63) / (2 * h)
        d = np.abs(fd - jac[:, :, t])
        bad += int((d > 2e-2 * max(1.0, np.abs(jac).max())).sum())
        total += d.size
    # LeakyReLU kinks inside +-h and the near-pi branch make isolated entries non-smooth
    assert bad <= 0.02 * total


def test_vposer_errors(vposer_params):
    from smplpp_b200 import api
    bad = dict(vposer_params)
    bad["decoder_net.3.weight"] = np.zeros((512, 511), f32)
    with pytest.raises(api.SmplppError, match="invalid dimension of decoder_net.3.weight"):
        api.VPoserDecoder(bad)
    with pytest.raises(api.SmplppError, match="Cannot find a JSON file!"):
        api.VPoserDecoder().loadParamsFromJson("/nonexistent/vposer_parameters.json")


@pytest.mark.parametrize("batch", [1, 3, 4, 257, 2048])
def test_tensor_core_jacobian_matches_ffma_kernel(vposer_gpu, batch):
    """The tcgen05 Jacobian (vposer_tc.cu: four frames per 128-lane tile, fp16 x 3 split precision) against the FFMA
    forward-mode kernel on the same latents, every frame and entry; ragged batches (not a multiple of 4) included.
    Tolerance: 1e-4 of the frame's largest entry, the north-star Jacobian bound; in practice ~6e-6."""
    from smplpp_b200 import capi
    z = np.random.default_rng(100 + batch).normal(size=(batch, 32)).astype(np.float32)
    out = {}
    for var in (301, 300):
        capi.check(capi.lib().smplpp_set_forward_variant(var))
        try:
            aa, jac = vposer_gpu.forward(z, jacobian=True)
            out[var] = (aa.clone(), jac.clone())
        finally:
            capi.check(capi.lib().smplpp_set_forward_variant(300))
    assert torch.equal(out[300][0], out[301][0])  # the forward pass is the same kernel
    assert bool(torch.isfinite(out[300][1]).all())
    scale = out[301][1].abs().amax(dim=(1, 2), keepdim=True)
    rel = ((out[300][1] - out[301][1]).abs() / scale).max()
    assert float(rel) < 1e-4
    assert float(rel) < 2e-5


@pytest.mark.parametrize("scales", [(8.0, 0.05, 3.0), (0.02, 6.0, 0.3), (1.0, 1.0, 40.0)])
def test_tensor_core_jacobian_weight_scales(vposer_params, scales):
    """The power-of-two operand scales of the tcgen05 Jacobian are derived from the weights at create time: decoders whose
    layers are much larger / smaller than the synthetic ones (x8, x0.02, x40 ...) must neither overflow fp16 nor lose the
    low parts to subnormals."""
    from smplpp_b200 import api, capi
    p = {k: np.array(v, dtype=np.float32, copy=True) for k, v in vposer_params.items()}
    p["decoder_net.0.weight"] *= scales[0]
    p["decoder_net.3.weight"] *= scales[1]
    p["decoder_net.5.weight"] *= scales[2]
    vp = api.VPoserDecoder(p, device="cuda:0")
    z = np.random.default_rng(7).normal(size=(64, 32)).astype(np.float32)
    out = {}
    for var in (301, 300):
        capi.check(capi.lib().smplpp_set_forward_variant(var))
        try:
            out[var] = vp.forward(z, jacobian=True)[1].clone()
        finally:
            capi.check(capi.lib().smplpp_set_forward_variant(300))
    assert bool(torch.isfinite(out[300]).all())
    scale = out[301].abs().amax(dim=(1, 2), keepdim=True).clamp_min(1e-30)
    assert float(((out[300] - out[301]).abs() / scale).max()) < 1e-4
