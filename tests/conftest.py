"""Shared fixtures.  `-m "not gpu"` covers the oracle, the golden vectors, host logic and the C-ABI exports;
`-m gpu` holds the parity tests proper (CUDA path vs oracle), all of which go through the C-ABI library."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# parity tolerances stated by BASELINE.json north_star
TOL_VERTEX_M = 1e-5  # vertices: max-abs metres
TOL_JACOBIAN_REL = 1e-4  # Jacobians: ||dJ||_max / ||J||_max per frame
TOL_RESIDUAL_M = 1e-4  # converged marker residual


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def params():
    from smplpp_b200 import synth
    return synth.make_smpl_params(0)


@pytest.fixture(scope="session")
def vposer_params():
    from smplpp_b200 import synth
    return synth.make_vposer_params(1)


@pytest.fixture(scope="session")
def oracle_model(params):
    from oracle import smpl_oracle
    return smpl_oracle.SmplModel.from_params(params)


@pytest.fixture(scope="session")
def oracle_vposer(vposer_params):
    from oracle import smpl_oracle
    return smpl_oracle.VPoserDecoder.from_params(vposer_params)


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(GOLDEN, "tester_kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_forward():
    return dict(np.load(os.path.join(GOLDEN, "ref_forward.npz")))


@pytest.fixture(scope="session")
def golden_vposer():
    return dict(np.load(os.path.join(GOLDEN, "ref_vposer.npz")))


@pytest.fixture(scope="session")
def golden_ik():
    return dict(np.load(os.path.join(GOLDEN, "ref_ik.npz")))


@pytest.fixture(scope="session")
def marker_tasks(params):
    from smplpp_b200 import synth
    return synth.make_marker_tasks(params)


@pytest.fixture(scope="session")
def smpl_gpu(params):
    """The product SMPL handle on cuda:0 (C-ABI library)."""
    from smplpp_b200 import api
    return api.SMPL(params, device="cuda:0")
