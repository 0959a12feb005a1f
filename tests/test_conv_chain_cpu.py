"""Host side of the Conv_AE inference path: every (transposed) convolution of the model becomes a dense matrix on the
host (models.Conv_AE._chains, BatchNorm2d folded).  The chain evaluated in numpy float64 must reproduce the reference
module's own eval outputs (tests/golden/conv_ae.npz, conv_shapes.npz) for all three block shapes the reference model
accepts - no device involved, so this also runs where there is no GPU."""
import numpy as np
import pytest
import torch

from conftest import randomise_bn2d, rel_max
from baler_b200.modules import models


def run_chain(layers, x):
    v = np.asarray(x, dtype=np.float64)
    for w, b, act in layers:
        v = v @ np.asarray(w, dtype=np.float64).T + np.asarray(b, dtype=np.float64)
        if act == "relu":
            v = np.maximum(v, 0.0)
        else:
            assert act == "none"
    return v


@pytest.mark.parametrize("fixture,tag,h,w,z_dim,conv_out", [("conv_ae.npz", "", 5, 5, 250, (32, 4, 1)),
                                                          ("conv_shapes.npz", "b36/", 3, 6, 9, (32, 2, 2)),
                                                          ("conv_shapes.npz", "b28/", 2, 8, 4, (32, 1, 4))])
def test_dense_chain_reproduces_the_reference_convolutions(golden, fixture, tag, h, w, z_dim, conv_out):
    g = golden(fixture)
    torch.manual_seed(0)
    m = models.Conv_AE(w, z_dim)
    m.load_state_dict(randomise_bn2d(m.state_dict()))
    enc, dec, shape = m._chains(h, w)
    assert tuple(shape) == conv_out
    assert [l[0].shape for l in enc] == [(8 * (h + 1) * (w - 2), h * w), (16 * (h + 1) * (w - 2), 8 * (h + 1) * (w - 2)),
                                         (128, 16 * (h + 1) * (w - 2)), (2000, 128), (z_dim, 2000)]
    x = g[tag + "blocks"].reshape(-1, h * w)
    z = run_chain(enc, x)
    # the reference computed in float32: its own rounding is the distance
    assert rel_max(z, g[tag + "latent_eval"]) <= 2e-6
    y = run_chain(dec, g[tag + "latent_eval"])
    assert rel_max(y, g[tag + "recon_eval"].reshape(-1, h * w)) <= 2e-6


def test_invalid_block_shapes_are_refused_like_upstream():
    torch.manual_seed(0)
    m = models.Conv_AE(5, 250)
    for h, w in ((50, 50), (4, 4), (6, 6)):
        with pytest.raises(RuntimeError, match="flattens"):
            m._chains(h, w)
