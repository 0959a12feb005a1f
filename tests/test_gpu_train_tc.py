"""The tensor-core (split16) training step, layer by layer against a float64 forward / backward of the same weights:
localises a wrong fragment layout or packing index to one layer pass.  Also: both arithmetic paths of bb_trainer agree,
the persistent epoch kernel equals the step-by-step API, and switching paths mid-run keeps the weights in sync."""
import numpy as np
import pytest
import torch

from conftest import rel_l2, rel_max, sub_sd
from oracle import baler_oracle as orc
from baler_b200 import engine
from baler_b200.modules import models

pytestmark = pytest.mark.gpu
NAMES = models.AE.names


def make_trainer(sd, max_batch=512, precision="auto"):
    tr = engine.Trainer([sd[n + ".weight"] for n in NAMES], [sd[n + ".bias"] for n in NAMES], 24, 15, max_batch)
    tr.set_precision(precision)
    return tr


def reference_pass(sd, x):
    """layer inputs X_l and pre-activation gradients dZ_l of the dense AE (oracle arithmetic, float64)"""
    acts, pre, h = [x], [], x
    for i, n in enumerate(NAMES):
        a = orc.linear(h, sd[n + ".weight"], sd[n + ".bias"])
        pre.append(a)
        h = a if i in (3, 7) else orc.leaky_relu(a)
        acts.append(h)
    d = 2.0 * (h - x) / x.shape[1]
    dz = [None] * 8
    for i in range(7, -1, -1):
        if i not in (3, 7):
            d = d * np.where(pre[i] > 0, 1.0, orc.LEAKY_SLOPE)
        dz[i] = d
        d = d @ sd[NAMES[i] + ".weight"]
    return acts, dz


@pytest.mark.parametrize("rows", [512, 448, 100, 16, 5])
def test_layer_by_layer(golden, rows):
    g = golden("ae_train.npz")
    sd0 = sub_sd(g, "sd0")
    x = g["x_norm"][:rows]
    tr = make_trainer(sd0)
    assert tr.precision == "split16"
    tr.step(torch.from_numpy(x).cuda(), engine.make_hyper(), phase=1)
    acts, dz = reference_pass(sd0, x.astype(np.float64))
    report = []
    for l in range(8):
        xl = tr.debug_layer(0, l, rows)
        assert xl.shape[0] == acts[l].shape[1] + 1
        report.append(("X%d" % l, rel_max(xl[:-1].T, acts[l]), float(np.abs(xl[-1] - 1).max())))
    for l in range(8):
        zl = tr.debug_layer(1, l, rows)
        report.append(("dZ%d" % l, rel_max(zl.T, dz[l]), 0.0))
    bad = [r for r in report if r[1] > 2e-6 or r[2] != 0.0]
    assert not bad, report
    assert not tr.range_flag()


def test_paths_agree_and_stay_in_sync(golden):
    g = golden("ae_train.npz")
    sd0 = sub_sd(g, "sd0")
    x = torch.from_numpy(g["x_norm"]).cuda()  # 2048 rows = 4 batches
    h = engine.make_hyper(lr=1e-3)
    a, b = make_trainer(sd0, precision="split16"), make_trainer(sd0, precision="fp32")
    assert (a.precision, b.precision) == ("split16", "fp32")
    for s in range(2):
        xb = x[s * 512:(s + 1) * 512].contiguous()
        a.step(xb, h)
        b.step(xb, h)
    pa, pb = a.params_view().cpu().numpy(), b.params_view().cpu().numpy()
    # Adam divides by |g|: where the gradient is small its relative rounding error (not its absolute one) moves the weight,
    # so two correct fp32-class implementations differ by ~1e-5 of max|w| after a few steps (the same bound
    # test_gpu_training.py holds both of them to against the float64 reference)
    assert rel_max(pa, pb) <= 2e-5, rel_max(pa, pb)
    assert abs(a.loss_accum.item() - b.loss_accum.item()) <= 1e-5 * b.loss_accum.item()
    # switch arithmetic mid-run in both directions: the derived weight copies of the other path are rebuilt from the
    # shared fp32 master parameters
    a.set_precision("fp32"); b.set_precision("split16")
    xb = x[2 * 512:3 * 512].contiguous()
    a.step(xb, h)
    b.step(xb, h)
    a.set_precision("split16"); b.set_precision("fp32")
    xb = x[3 * 512:4 * 512].contiguous()
    a.step(xb, h)
    b.step(xb, h)
    pa, pb = a.params_view().cpu().numpy(), b.params_view().cpu().numpy()
    assert rel_max(pa, pb) <= 4e-5, rel_max(pa, pb)
    assert abs(a.loss_accum.item() - b.loss_accum.item()) <= 1e-5 * b.loss_accum.item()


def test_epoch_kernel_equals_step_api(golden):
    """one persistent launch over 4 batches (the last one ragged) == four bb_trainer_step calls, bit for bit"""
    g = golden("ae_train.npz")
    sd0 = sub_sd(g, "sd0")
    x = torch.from_numpy(g["x_norm"][:3 * 512 + 200]).cuda()
    assert x.shape[0] == 1736
    h = engine.make_hyper(lr=1e-3)
    a, b = make_trainer(sd0), make_trainer(sd0)
    loss = a.epoch(x, 512, h)
    for r0 in range(0, x.shape[0], 512):
        b.step(x[r0:r0 + 512].contiguous(), h)
    assert torch.equal(a.params_view(), b.params_view())
    assert abs(loss - b.loss_accum.item() / 4) <= 1e-12 * loss
    # validation pass (forward only) on both arithmetic paths
    va = a.validate(x, 512)
    a.set_precision("fp32")
    vb = a.validate(x, 512)
    assert abs(va - vb) <= 1e-5 * vb, (va, vb)
    # a second epoch continues the Adam step count
    a.set_precision("split16")
    a.epoch(x, 512, h)
    for r0 in range(0, x.shape[0], 512):
        b.step(x[r0:r0 + 512].contiguous(), h)
    assert torch.equal(a.params_view(), b.params_view())


def test_large_batch_loops_tiles_and_panels(golden):
    """batch 2048 > 512: several 16-row tiles per CTA in phase 1, several 512-row panel passes in phase 2"""
    g = golden("ae_train.npz")
    sd0 = sub_sd(g, "sd0")
    x = g["x_norm"][:2048]
    loss, _, _, grads = orc.ae_loss_and_grads(sd0, x.astype(np.float64))
    tr = make_trainer(sd0, max_batch=2048)
    tr.step(torch.from_numpy(x).cuda(), engine.make_hyper(), phase=1)
    got = tr.grads_view().cpu().numpy()
    ref = np.concatenate([np.concatenate([grads[n + ".weight"].ravel(), grads[n + ".bias"].ravel()]) for n in NAMES])
    assert abs(got[-1] - loss) <= 1e-5 * loss
    assert rel_max(got[:-1], ref) <= 1e-5 and rel_l2(got[:-1], ref) <= 1e-5
