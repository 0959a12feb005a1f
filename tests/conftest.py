import os
import sys

import numpy as np
import torch
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    # some fixtures were saved Fortran-ordered (np.apply_along_axis output): hand out C-contiguous arrays
    return {k: np.asarray(v, order="C") for k, v in np.load(os.path.join(GOLDEN, name), allow_pickle=False).items()}


def sub_sd(d, prefix):
    """state dict stored under '<prefix>/<key>' in a golden npz"""
    p = prefix + "/"
    return {k[len(p):]: v for k, v in d.items() if k.startswith(p)}


def rel_max(a, ref):
    """max|a - ref| / max|ref| - the tolerance norm of SURVEY.md 8(d)"""
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-300))


def rel_l2(a, ref):
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.linalg.norm((a - ref).ravel()) / max(np.linalg.norm(ref.ravel()), 1e-300))


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]

    return get


def randomise_bn2d(sd, seed=2):
    """same draws as oracle/gen_golden.py::randomise_bn2d, in module order"""
    g = torch.Generator().manual_seed(seed)
    for name in ("q_z_conv.3", "p_x_conv.1", "p_x_conv.4"):
        n = sd[name + ".weight"].shape
        sd[name + ".weight"] = 0.5 + torch.rand(n, generator=g)
        sd[name + ".bias"] = 0.2 * torch.randn(n, generator=g)
        sd[name + ".running_mean"] = 0.1 * torch.randn(n, generator=g)
        sd[name + ".running_var"] = 0.5 + torch.rand(n, generator=g)
    return sd
