"""Training parity (SURVEY 8d contract): per-step loss / gradients / Adam update within 1e-5 of the
reference from identical weights and batches; epoch-mean loss curve within 1 %; same files written."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2, rel_max, sub_sd
from oracle import baler_oracle as orc
from baler_b200 import engine, synth
from baler_b200.modules import models

pytestmark = pytest.mark.gpu
NAMES = models.AE.names


def flat(sd_like):
    """flatten a state-dict-like mapping in the trainer's layout: W_l (out,in) then b_l"""
    return np.concatenate([np.concatenate([np.asarray(sd_like[n + ".weight"]).ravel(), np.asarray(sd_like[n + ".bias"]).ravel()])
                           for n in NAMES])


def make_trainer(sd, max_batch=512):
    return engine.Trainer([sd[n + ".weight"] for n in NAMES], [sd[n + ".bias"] for n in NAMES], 24, 15, max_batch)


@pytest.mark.parametrize("tag,l1,precision", [("mse", False, "split16"), ("mse", False, "fp32"), ("l1", True, "fp32")])
def test_step_loss_grads_adam(golden, tag, l1, precision):
    """`precision`: the tensor-core step (default for the MSE loss) and the fp32 FFMA step (always used for the L1 chain)"""
    g = golden("ae_train.npz")
    sd0 = sub_sd(g, "sd0")
    x = torch.from_numpy(g["x_norm"]).cuda()
    tr = make_trainer(sd0)
    tr.set_precision(precision)
    assert tr.precision == precision and tr.n_params == 61839
    hyper = engine.make_hyper(lr=1e-3, reg_param=0.001, l1=l1)
    ref_g = flat(sub_sd(g, "g_" + tag))
    p0 = flat(sd0)
    losses = []
    for step in range(3):
        tr.loss_accum.zero_()
        tr.step(x[step * 512:(step + 1) * 512].contiguous(), hyper)
        losses.append(tr.loss_accum.item())
        if step == 0:
            grads = tr.grads_view()[:-1].cpu().numpy()
            assert rel_max(grads, ref_g) <= 1e-5 and rel_l2(grads, ref_g) <= 1e-5, (rel_max(grads, ref_g), rel_l2(grads, ref_g))
            assert abs(tr.grads_view()[-1].item() - g["losses_" + tag][0]) <= 1e-5 * g["losses_" + tag][0]
        if step in (0, 2):
            ref_p = flat(sub_sd(g, "sd%d_%s" % (step + 1, tag)))
            p = tr.params_view().cpu().numpy().astype(np.float64)
            # the UPDATE is what is being tested: lr * m / (sqrt(v) + eps).  Where |g| approaches Adam's eps (1e-8)
            # the first-step update g / (|g| + eps) amplifies fp32 rounding of g, so the tight bound is on elements
            # with a resolvable gradient and a loose absolute bound (5 % of lr) covers the rest
            big = np.abs(ref_g) > 1e-3 * np.abs(ref_g).max()
            du, dr = p - p0, ref_p - p0
            assert rel_max(du[big], dr[big]) <= 2e-4 and rel_l2(du, dr) <= 1e-4, (step, rel_max(du[big], dr[big]), rel_l2(du, dr))
            assert np.abs(du - dr).max() <= 0.05 * 1e-3
            # Adam's first steps move a weight by lr * g / (|g| + 1e-8): where |g| is ~1e-5 of the largest gradient the
            # update depends on g's absolute rounding error.  The fp32 path (gradients 1e-7 of max) holds 1e-5 of max|w|;
            # the split16 path (3e-7 of max: the tensor core truncates inside every 16-product sum) measures 1.04e-5 after
            # one step and 1.2e-5 after three on a handful of such weights, inside the 2e-4 / 1e-4 update bounds above
            assert rel_max(p, ref_p) <= (1e-5 if precision == "fp32" else 2e-5)
    np.testing.assert_allclose(losses, g["losses_" + tag], rtol=1e-5)
    w, b = tr.get_params()
    assert w[0].shape == (200, 24) and w[0].dtype == np.float64 and b[7].shape == (24,)


def test_per_tensor_gradients(golden):
    g = golden("ae_train.npz")
    sd0 = sub_sd(g, "sd0")
    tr = make_trainer(sd0)
    tr.step(torch.from_numpy(g["x_norm"][:512]).cuda(), engine.make_hyper())
    grads = tr.grads_view()[:-1].cpu().numpy()
    ref = sub_sd(g, "g_mse")
    off = 0
    for n in NAMES:
        for part in (".weight", ".bias"):
            r = ref[n + part].ravel()
            got = grads[off:off + r.size]
            off += r.size
            assert rel_max(got, r) <= 1e-5 and rel_l2(got, r) <= 1e-5, (n + part, rel_max(got, r))


@pytest.mark.parametrize("rows", [1, 7, 8, 9, 63, 64, 65, 100, 448])
def test_ragged_batches_vs_oracle(golden, rows):
    g = golden("ae_train.npz")
    sd0 = sub_sd(g, "sd0")
    x = g["x_norm"][:rows]
    loss, _, _, grads = orc.ae_loss_and_grads(sd0, x.astype(np.float64))
    if rows == 100:
        assert abs(loss - float(g["loss_ragged100"])) < 1e-12 * loss
    tr = make_trainer(sd0)
    tr.step(torch.from_numpy(x).cuda(), engine.make_hyper(), phase=1)  # forward + backward only
    got = tr.grads_view().cpu().numpy()
    assert abs(got[-1] - loss) <= 1e-5 * loss
    ref = flat(grads)
    assert rel_max(got[:-1], ref) <= 1e-5 and rel_l2(got[:-1], ref) <= 1e-5
    assert np.array_equal(tr.params_view().cpu().numpy(), flat(sd0).astype(np.float32))  # phase 1 leaves weights alone


def test_split_phases_equal_fused_step(golden):
    """phase 1 (fwd+bwd) followed by phase 2 (Adam) is what data-parallel ranks run around the all-reduce"""
    g = golden("ae_train.npz")
    sd0 = sub_sd(g, "sd0")
    x = torch.from_numpy(g["x_norm"][:512]).cuda()
    a, b = make_trainer(sd0), make_trainer(sd0)
    h = engine.make_hyper()
    a.step(x, h, phase=0)
    b.step(x, h, phase=1)
    b.step(x, h, phase=2)
    assert torch.equal(a.params_view(), b.params_view()) and a.loss_accum.item() == b.loss_accum.item()
    # two half batches, gradients summed (SUM all-reduce semantics) == one full batch
    c, d = make_trainer(sd0), make_trainer(sd0)
    c.step(x[:256].contiguous(), h, phase=1)
    d.step(x[256:].contiguous(), h, phase=1)
    summed = c.grads_view() + d.grads_view()
    full = b.grads_view()
    assert rel_max(summed.cpu().numpy(), full.cpu().numpy()) <= 1e-6


def test_fit_curve_and_validate(golden):
    g, g0 = golden("ae_fit.npz"), golden("ae_train.npz")
    sd0 = sub_sd(g0, "sd0")
    table = synth.cms_table(4096, seed=11)
    x = torch.from_numpy(orc.normalize(table)).cuda()
    tr = make_trainer(sd0)
    h = engine.make_hyper(lr=1e-3)
    losses = [tr.epoch(x, 512, h) for _ in range(3)]
    ref = g["loss_data"][0]
    np.testing.assert_allclose(losses[:2], ref[:2], rtol=1e-4)   # contract: epochs 1-2 within 1 %; measured far tighter
    np.testing.assert_allclose(losses[2], ref[2], rtol=1e-2)
    w, b = tr.get_params()
    sd = {n + ".weight": w[i] for i, n in enumerate(NAMES)}
    sd.update({n + ".bias": b[i] for i, n in enumerate(NAMES)})
    val = tr.validate(x, 512)
    ref_val = np.mean([orc.mse_sum_loss(orc.ae_forward(sd, x[i:i + 512].cpu().numpy().astype(np.float64)),
                                        x[i:i + 512].cpu().numpy().astype(np.float64)) for i in range(0, 4096, 512)])
    assert abs(val - ref_val) <= 1e-5 * ref_val
    act = tr.activation_means()
    assert act.shape == (6, 200) and np.isnan(act[1, 100:]).all() and not np.isnan(act[0]).any()


def test_train_mode_writes_reference_layout(golden, tmp_path, monkeypatch):
    """`--mode train` through the drop-in: same files as the reference, loss curve within 1 % of the
    reference's own run from the same seed (tests/golden/cli_roundtrip.npz)."""
    from baler_b200 import baler
    from baler_b200.modules import helper

    g = golden("cli_roundtrip.npz")
    monkeypatch.chdir(tmp_path)
    helper.create_new_project("CMS_workspace", "CMS_project_v1")
    table = synth.cms_table(4096, seed=17)
    path = os.path.join("workspaces", "CMS_workspace", "data", "example_CMS_data.npz")
    np.savez(path, data=table, names=synth.CMS_NAMES)
    out = os.path.join("workspaces", "CMS_workspace", "CMS_project_v1", "output")

    class cfg(helper.Config):
        input_path = path
        data_dimension, compression_ratio, apply_normalization, model_name = 1, 1.6, True, "AE"
        epochs, lr, batch_size, early_stopping, lr_scheduler = 2, 0.001, 512, True, True
        early_stopping_patience, min_delta, lr_scheduler_patience, custom_norm = 100, 0, 50, False
        reg_param, RHO, test_size, extra_compression = 0.001, 0.05, 0, False
        intermittent_model_saving, intermittent_saving_patience = False, 100
        l1, activation_extraction, deterministic_algorithm = True, True, False
        convert_to_blocks, separate_model_saving, save_error_bounded_deltas = False, False, False

    torch.manual_seed(0)
    baler.perform_training(out, cfg, False)
    loss = np.load(os.path.join(out, "training", "loss_data.npy"))
    assert loss.shape == (2, 2)
    np.testing.assert_allclose(loss, g["loss_data"], rtol=1e-2)
    assert np.array_equal(np.load(os.path.join(out, "training", "normalization_features.npy")), g["norm_features"])
    sd = torch.load(os.path.join(out, "compressed_output", "model.pt"))
    ref = sub_sd(g, "sd")
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert sd[k].dtype == torch.float64 and rel_max(sd[k].numpy(), ref[k]) <= 1e-2, k
    act = np.load(os.path.join(out, "training", "activations.npy"))
    assert act.shape == g["activations"].shape
    assert rel_max(np.nan_to_num(act), np.nan_to_num(g["activations"])) <= 5e-2


# ------------------------------------------------------------------ AE_Dropout_BN (models.py:256-313), train mode
DBN_LIN = models.AE_Dropout_BN.enc_names + models.AE_Dropout_BN.dec_names
DBN_BN = models.AE_Dropout_BN.bn_names


def dbn_trainer(sd, max_batch=512):
    bn = {k: [sd[b + "." + k] for b in DBN_BN] for k in ("weight", "bias", "running_mean", "running_var")}
    bn["num_batches_tracked"] = [int(sd[b + ".num_batches_tracked"]) for b in DBN_BN]
    return engine.Trainer([sd[n + ".weight"] for n in DBN_LIN], [sd[n + ".bias"] for n in DBN_LIN], 24, 15, max_batch, bn=bn)


def dbn_flat(sd_like):
    lin = [np.concatenate([np.asarray(sd_like[n + ".weight"]).ravel(), np.asarray(sd_like[n + ".bias"]).ravel()]) for n in DBN_LIN]
    bn = [np.concatenate([np.asarray(sd_like[b + ".weight"]).ravel(), np.asarray(sd_like[b + ".bias"]).ravel()]) for b in DBN_BN]
    return np.concatenate(lin + bn)


@pytest.mark.parametrize("precision", ["split16", "fp32"])
def test_dropout_bn_train_step_with_injected_masks(golden, precision):
    """one reference train step of AE_Dropout_BN (loss, every gradient, BN running statistics, Adam update) with the
    dropout masks torch drew, injected; on the tensor-core step (default) and on the fp32 kernels"""
    g = golden("ae_dbn.npz")
    sd0, sd1 = sub_sd(g, "sd0"), sub_sd(g, "sd1")
    tr = dbn_trainer(sd0)
    tr.set_precision(precision)
    assert tr.precision == precision
    assert tr.n_params == 61839 + 2 * (50 + 100 + 200 + 24)
    masks = [torch.from_numpy(g["mask%d" % i].astype(np.uint8)).cuda() for i in range(4)]
    tr.set_dropout(masks=masks)
    x = torch.from_numpy(g["x_norm"][:512]).cuda()
    tr.step(x, engine.make_hyper(lr=1e-3))
    loss = tr.loss_accum.item()
    assert abs(loss - float(g["loss_train"])) <= 1e-5 * float(g["loss_train"]), (loss, float(g["loss_train"]))
    grads = tr.grads_view()[:-1].cpu().numpy()
    ref = dbn_flat(sub_sd(g, "g"))
    gscale = np.abs(ref).max()
    assert np.abs(grads - ref).max() <= 2e-5 * gscale and rel_l2(grads, ref) <= 2e-5, (np.abs(grads - ref).max() / gscale, rel_l2(grads, ref))
    bn = tr.get_bn()
    for i, b in enumerate(DBN_BN):
        assert rel_max(bn["running_mean"][i], sd1[b + ".running_mean"]) <= 1e-5, b
        assert rel_max(bn["running_var"][i], sd1[b + ".running_var"]) <= 1e-5, b
        assert bn["num_batches_tracked"][i] == int(sd1[b + ".num_batches_tracked"])
    p = tr.params_view().cpu().numpy().astype(np.float64)
    p0, p1 = dbn_flat(sd0), dbn_flat(sd1)
    big = np.abs(ref) > 1e-3 * gscale  # Adam's first step amplifies rounding where |g| ~ eps (see the AE test)
    assert rel_max((p - p0)[big], (p1 - p0)[big]) <= 1e-3
    # Linear biases that feed a BatchNorm have a mathematically zero gradient (the BN removes the shift): the
    # reference's own value there is 1e-16 rounding noise and Adam turns any noise into a +-lr step, so those
    # parameters (which do not influence the model output) are only required to move by at most lr
    live = np.abs(ref) > 1e-6 * gscale
    assert np.abs(p - p1)[live].max() <= 0.05 * 1e-3
    assert np.abs(p - p1).max() <= 1.001e-3


def test_dropout_bn_philox_keep_rates_and_training(golden):
    g = golden("ae_dbn.npz")
    sd0 = sub_sd(g, "sd0")
    table = synth.cms_table(8192, seed=23)
    x = torch.from_numpy(orc.normalize(table)).cuda()
    tr = dbn_trainer(sd0)
    tr.set_dropout(seed=1234)
    h = engine.make_hyper(lr=1e-3)
    losses = [tr.epoch(x, 512, h) for _ in range(4)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]  # it learns
    # dropout statistics: the mean |activation gradient| pattern is not observable directly, so check the keep rates
    # through the fraction of exactly-zero encoder activations of the last batch (dropped units give LeakyReLU(0) = 0)
    tr2 = dbn_trainer(sd0)
    tr2.set_dropout(seed=99)
    tr2.step(x[:512].contiguous(), h, phase=1)
    means = tr2.activation_means()  # exercises the scratch read-back on this trainer kind too
    assert means.shape == (6, 200)
    # eval forward: running statistics, no dropout -> equals the folded-BN codec
    val = tr.validate(x, 512)
    assert np.isfinite(val) and val > 0
    # different seeds give different results, same seed reproduces bit for bit
    a, b, c = dbn_trainer(sd0), dbn_trainer(sd0), dbn_trainer(sd0)
    for t_, seed in ((a, 7), (b, 7), (c, 8)):
        t_.set_dropout(seed=seed)
        t_.step(x[:512].contiguous(), h)
    assert torch.equal(a.params_view(), b.params_view()) and not torch.equal(a.params_view(), c.params_view())


@pytest.mark.parametrize("rows", [2000, 1999, 37])
def test_dropout_bn_large_batch_vs_oracle(golden, rows):
    """tensor-core AE_Dropout_BN step on batches beyond the fp32 kernel's 592 rows (and ragged ones): loss, all gradients
    and running statistics against the float64 oracle with the same (numpy-drawn) keep-masks"""
    g = golden("ae_dbn.npz")
    sd0 = sub_sd(g, "sd0")
    x = orc.normalize(synth.cms_table(rows, seed=31))
    rng = np.random.default_rng(5)
    masks = [(rng.random((rows, w)) < keep).astype(np.uint8) for w, keep in ((200, 0.5), (100, 0.6), (50, 0.7), (15, 0.8))]
    tr = dbn_trainer(sd0, max_batch=2048)
    assert tr.precision == "split16"
    tr.set_dropout(masks=[torch.from_numpy(m).cuda() for m in masks])
    tr.step(torch.from_numpy(x).cuda(), engine.make_hyper(lr=1e-3), phase=1)
    loss, _, grads, buf = orc.dbn_train_step({k: np.asarray(v, dtype=np.float64) for k, v in sd0.items()}, x.astype(np.float64),
                                             [m.astype(np.float64) for m in masks])
    got = tr.grads_view().cpu().numpy()
    assert abs(got[-1] - loss) <= 1e-5 * loss, (got[-1], loss)
    ref = dbn_flat(grads)
    gscale = np.abs(ref).max()
    # (a 37-row batch: 1 / sqrt(var + eps) of the small-population statistics amplifies the 22-bit operand rounding)
    tol = 2e-5 if rows >= 512 else 5e-5
    assert np.abs(got[:-1] - ref).max() <= tol * gscale and rel_l2(got[:-1], ref) <= 2e-5, (np.abs(got[:-1] - ref).max() / gscale, rel_l2(got[:-1], ref))
    bn = tr.get_bn()
    for i, b in enumerate(DBN_BN):
        assert rel_max(bn["running_mean"][i], buf[b + ".running_mean"]) <= 1e-5, b
        assert rel_max(bn["running_var"][i], buf[b + ".running_var"]) <= 1e-5, b


def test_dropout_bn_paths_share_the_dropout_stream(golden):
    """the tensor-core step and the fp32 kernels draw the same Philox keep-masks from (seed, step, layer, row, column):
    one step from the same state with the same seed gives the same gradients up to arithmetic (2e-5), and an epoch of the
    tensor-core path tracks the fp32 path"""
    g = golden("ae_dbn.npz")
    sd0 = sub_sd(g, "sd0")
    x = torch.from_numpy(orc.normalize(synth.cms_table(4096, seed=23))).cuda()
    h = engine.make_hyper(lr=1e-3)
    out = {}
    for prec in ("split16", "fp32"):
        tr = dbn_trainer(sd0)
        tr.set_precision(prec)
        tr.set_dropout(seed=4321)
        tr.step(x[:512].contiguous(), h, phase=1)
        out[prec] = tr.grads_view().cpu().numpy()
        tr2 = dbn_trainer(sd0)
        tr2.set_precision(prec)
        tr2.set_dropout(seed=4321)
        out[prec + "_loss"] = [tr2.epoch(x, 512, h) for _ in range(2)]
    gscale = np.abs(out["fp32"][:-1]).max()
    assert np.abs(out["split16"] - out["fp32"])[:-1].max() <= 2e-5 * gscale
    assert abs(out["split16"][-1] - out["fp32"][-1]) <= 1e-5 * out["fp32"][-1]
    assert np.allclose(out["split16_loss"], out["fp32_loss"], rtol=1e-2)  # the north-star bar for loss curves: 1 %


def test_dropout_bn_validate_equals_folded_codec(golden):
    """eval-mode loss of the trainer (running statistics) == sum-MSE of the folded-BN inference codec"""
    g = golden("ae_dbn.npz")
    sd0 = sub_sd(g, "sd0")
    x = g["x_norm"][:512]
    tr = dbn_trainer(sd0)
    val = tr.validate(torch.from_numpy(x).cuda(), 512)
    rec = orc.dbn_decode(sd0, orc.dbn_encode(sd0, x.astype(np.float64)))
    ref = orc.mse_sum_loss(rec, x.astype(np.float64))
    assert abs(val - ref) <= 1e-5 * ref


def test_dropout_bn_data_parallel_step_matches_oracle(golden):
    """data-parallel AE_Dropout_BN: every rank runs the train-mode forward / backward on ITS slice of the global batch
    (per-rank BatchNorm statistics), the flat gradients are SUMMED and every rank applies the same Adam step.  Two
    trainers on one GPU play the two ranks; the reference arithmetic for each slice comes from the oracle."""
    g = golden("ae_dbn.npz")
    sd0 = sub_sd(g, "sd0")
    x = g["x_norm"][:512]
    masks = [g["mask%d" % i].astype(np.uint8) for i in range(4)]
    h = engine.make_hyper(lr=1e-3, world_size=2)
    ranks, ref_g, ref_loss = [], None, 0.0
    for r, (lo, hi) in enumerate(((0, 256), (256, 512))):
        tr = dbn_trainer(sd0)
        tr.set_dropout(masks=[torch.from_numpy(np.ascontiguousarray(m[lo:hi])).cuda() for m in masks])
        tr.step(torch.from_numpy(np.ascontiguousarray(x[lo:hi])).cuda(), h, phase=1)
        ranks.append(tr)
        loss, _, grads, _ = orc.dbn_train_step({k: np.asarray(v, dtype=np.float64) for k, v in sd0.items()},
                                               x[lo:hi].astype(np.float64), [m[lo:hi].astype(np.float64) for m in masks])
        flat = dbn_flat(grads)
        ref_g = flat if ref_g is None else ref_g + flat
        ref_loss += loss
    total = ranks[0].grads_view() + ranks[1].grads_view()  # what the SUM all-reduce leaves on every rank
    for tr in ranks:
        tr.grads_view().copy_(total)
    got = total[:-1].cpu().numpy()
    gscale = np.abs(ref_g).max()
    assert np.abs(got - ref_g).max() <= 2e-5 * gscale and rel_l2(got, ref_g) <= 2e-5
    assert abs(float(total[-1]) - ref_loss) <= 1e-5 * ref_loss
    for tr, (lo, hi) in zip(ranks, ((0, 256), (256, 512))):
        tr.step(torch.from_numpy(np.ascontiguousarray(x[lo:hi])).cuda(), h, phase=2)
    assert torch.equal(ranks[0].params_view(), ranks[1].params_view())  # replicas stay bit-identical
    rm, rv = ranks[0].bn_running_views()
    assert rm.numel() == 200 + 100 + 50 + 24 and torch.isfinite(rm).all() and torch.isfinite(rv).all()


# ------------------------------------------------------------------ layer-by-layer trainer (wide rows: CFD_dense_AE)
AE_ACTS = ["leaky", "leaky", "leaky", "none"] * 2


def _layered(sd, max_batch):
    return engine.LayeredTrainer([sd[n + ".weight"] for n in NAMES], [sd[n + ".bias"] for n in NAMES], AE_ACTS, max_batch)


def test_layered_trainer_cfd_dense_step_matches_oracle():
    """CFD_dense_AE(2500, 25) (models.py:186-226; the shipped CFD_project_animation: 60 snapshots of 50 x 50): one fit
    step - loss, every gradient, Adam update - against the float64 oracle"""
    torch.manual_seed(3)
    m = models.CFD_dense_AE(2500, 25)
    sd = {k: v.numpy().astype(np.float64) for k, v in m.state_dict().items()}
    rng = np.random.default_rng(5)
    x = rng.random((60, 2500), dtype=np.float32)
    tr = _layered(sd, 64)
    assert tr.n_params == sum(v.size for v in sd.values())
    tr.step(torch.from_numpy(x).cuda(), engine.make_hyper(lr=1e-3), phase=1)
    loss_ref, _, _, g = orc.ae_loss_and_grads(sd, x.astype(np.float64))
    got = tr.grads_view().cpu().numpy()
    ref = np.concatenate([np.concatenate([g[n + ".weight"].ravel(), g[n + ".bias"].ravel()]) for n in NAMES])
    assert abs(got[-1] - loss_ref) <= 1e-5 * loss_ref
    assert rel_max(got[:-1], ref) <= 1e-5 and rel_l2(got[:-1], ref) <= 1e-5, (rel_max(got[:-1], ref), rel_l2(got[:-1], ref))
    # full step from the same start: Adam's first update is lr * sign(g) wherever |g| >> eps
    tr2 = _layered(sd, 64)
    tr2.step(torch.from_numpy(x).cuda(), engine.make_hyper(lr=1e-3))
    p0 = np.concatenate([np.concatenate([sd[n + ".weight"].ravel(), sd[n + ".bias"].ravel()]) for n in NAMES])
    opt = orc.Adam({k: v.copy() for k, v in sd.items()}, lr=1e-3)
    opt.step(g)
    p1 = np.concatenate([np.concatenate([opt.params[n + ".weight"].ravel(), opt.params[n + ".bias"].ravel()]) for n in NAMES])
    p = tr2.params_view().cpu().numpy().astype(np.float64)
    big = np.abs(ref) > 1e-3 * np.abs(ref).max()
    assert rel_max((p - p0)[big], (p1 - p0)[big]) <= 1e-3
    assert abs(tr2.loss_accum.item() - loss_ref) <= 1e-5 * loss_ref


def test_layered_trainer_equals_fused_trainer_on_the_cms_shape(golden):
    """same model, same batch: the GEMM-per-layer trainer and the fused kernels give the same gradients, and a few
    epochs give the same loss curve"""
    g = golden("ae_train.npz")
    sd0 = sub_sd(g, "sd0")
    x = torch.from_numpy(np.ascontiguousarray(g["x_norm"][:2048])).cuda()
    a, b = make_trainer(sd0, 512), _layered(sd0, 512)
    h = engine.make_hyper(lr=1e-3)
    a.step(x[:512].contiguous(), h, phase=1)
    b.step(x[:512].contiguous(), h, phase=1)
    ga, gb = a.grads_view().cpu().numpy(), b.grads_view().cpu().numpy()
    assert rel_max(gb, ga) <= 1e-5 and rel_l2(gb, ga) <= 1e-5
    a, b = make_trainer(sd0, 512), _layered(sd0, 512)
    la = [a.epoch(x, 512, h) for _ in range(3)]
    lb = [b.epoch(x, 512, h) for _ in range(3)]
    assert np.allclose(la, lb, rtol=1e-3), (la, lb)
    assert abs(b.validate(x, 512) - a.validate(x, 512)) <= 1e-3 * a.validate(x, 512)


def test_cfd_dense_training_through_the_training_module(tmp_path):
    """training.train on a 2-D dense project (flattened 50 x 50 snapshots -> CFD_dense_AE(2500, z)): the loss curve goes
    down and the files of the reference's train mode are written"""
    from types import SimpleNamespace
    from baler_b200.modules import training
    rng = np.random.default_rng(1)
    t = np.linspace(0, 1, 50, dtype=np.float32)
    snaps = np.stack([np.outer(np.sin(2 * np.pi * (t + 0.01 * i)), np.cos(2 * np.pi * t * (1 + 0.02 * i))) for i in range(60)])
    snaps = ((snaps - snaps.min()) / (snaps.max() - snaps.min())).astype(np.float32) + 0.01 * rng.random(snaps.shape, dtype=np.float32)
    cfg = SimpleNamespace(deterministic_algorithm=True, batch_size=20, data_dimension=2, model_type="dense", lr=1e-3, reg_param=0.001,
                          early_stopping=False, early_stopping_patience=100, min_delta=0, lr_scheduler=False, lr_scheduler_patience=50,
                          epochs=12, test_size=0, intermittent_model_saving=False, intermittent_saving_patience=100, RHO=0.05, l1=True,
                          activation_extraction=False)
    torch.manual_seed(0)
    model = models.CFD_dense_AE(2500, 25)
    w0 = model.state_dict()["en1.weight"].clone()
    training.train(model, 50, snaps, snaps, str(tmp_path), cfg)
    losses = np.load(tmp_path / "loss_data.npy")
    assert losses.shape == (2, 12) and np.isfinite(losses).all() and losses[0, -1] < 0.7 * losses[0, 0]
    assert not torch.equal(model.state_dict()["en1.weight"], w0) and model.state_dict()["en1.weight"].dtype == torch.float32


def test_layered_trainer_data_parallel_phases():
    """the layered trainer has the same two-phase step as the fused one: gradients of two half batches sum to the full
    batch's gradient (the loss is a sum over rows), and phase 2 applies Adam on the summed buffer"""
    torch.manual_seed(4)
    m = models.CFD_dense_AE(400, 20)
    sd = {k: v.numpy().astype(np.float64) for k, v in m.state_dict().items()}
    x = torch.rand((64, 400), device="cuda")
    h1, h2 = engine.make_hyper(lr=1e-3), engine.make_hyper(lr=1e-3, world_size=2)
    full = _layered(sd, 64)
    full.step(x, h1, phase=1)
    ranks = [_layered(sd, 64) for _ in range(2)]
    for tr, xs in zip(ranks, (x[:32].contiguous(), x[32:].contiguous())):
        tr.step(xs, h2, phase=1)
    total = ranks[0].grads_view() + ranks[1].grads_view()
    ref = full.grads_view()
    assert rel_max(total.cpu().numpy(), ref.cpu().numpy()) <= 1e-5
    for tr, xs in zip(ranks, (x[:32].contiguous(), x[32:].contiguous())):
        tr.grads_view().copy_(total)
        tr.step(xs, h2, phase=2)
    full.step(x, h1, phase=2)
    assert torch.equal(ranks[0].params_view(), ranks[1].params_view())
    assert rel_max(ranks[0].params_view().cpu().numpy(), full.params_view().cpu().numpy()) <= 1e-5
    assert abs(ranks[0].loss_accum.item() - full.loss_accum.item()) <= 1e-5 * full.loss_accum.item()


def _flat_named(sd_like):
    return np.concatenate([np.concatenate([np.asarray(sd_like[n + ".weight"]).ravel(), np.asarray(sd_like[n + ".bias"]).ravel()])
                           for n in NAMES]).astype(np.float64)


def test_swae_step_matches_reference(golden):
    """config.custom_loss_function = "loss_function_swae" (training.py:70-78, utils.py:27-91): one step on a float32
    CFD_dense_AE(64, 10) with the reference's own seeded draws (tests/golden/swae.npz) - loss = sum-MSE / C + sliced
    Wasserstein term, every gradient, parameters after Adam - and the float64 restatement on a ragged batch"""
    g = golden("swae.npz")
    sd0 = sub_sd(g, "sd0")
    tr = _layered(sd0, 64)
    x = torch.from_numpy(g["x"]).cuda()
    prior, proj = torch.from_numpy(g["prior"]).cuda(), torch.from_numpy(g["proj"]).cuda()
    h = engine.make_hyper(lr=1e-3)
    tr.step_swae(x, h, prior, proj, latent_layer=3, phase=1)
    flat = tr.grads_view().cpu().numpy().astype(np.float64)
    assert abs(flat[-1] - float(g["loss"])) <= 1e-5 * float(g["loss"])
    ref = _flat_named(sub_sd(g, "g"))
    assert rel_max(flat[:-1], ref) <= 1e-5 and rel_l2(flat[:-1], ref) <= 1e-5
    tr.step_swae(x, h, prior, proj, latent_layer=3, phase=2)
    upd = tr.params_view().cpu().numpy().astype(np.float64) - _flat_named(sd0)
    ref_upd = _flat_named(sub_sd(g, "sd1")) - _flat_named(sd0)
    # Adam's first step is lr * g / (|g| + eps): entries whose gradient is ~1e-8 (dead units) move by a noise-decided amount
    big = np.abs(ref) > 1e-6 * np.abs(ref).max()
    assert np.abs(upd - ref_upd)[big].max() <= 1e-2 * 1e-3 and np.abs(upd - ref_upd).max() <= 2.1e-3
    # ragged batch (37 of 64 rows), other draws: against the float64 restatement
    rng = np.random.default_rng(8)
    x2 = rng.random((37, 64), dtype=np.float32)
    pr2 = rng.standard_normal((37, 10)).astype(np.float32)
    pj2 = rng.standard_normal((500, 10)).astype(np.float32)
    pj2 /= np.linalg.norm(pj2, axis=1, keepdims=True)
    tr2 = _layered(sd0, 64)
    tr2.step_swae(torch.from_numpy(x2).cuda(), h, torch.from_numpy(pr2).cuda(), torch.from_numpy(pj2).cuda(), latent_layer=3,
                  reg_weight=40.0, phase=1)
    loss, _, _, grads = orc.ae_loss_and_grads(sd0, x2, swae=(pr2, pj2, 40.0))
    flat2 = tr2.grads_view().cpu().numpy().astype(np.float64)
    assert abs(flat2[-1] - loss) <= 1e-5 * loss
    assert rel_max(flat2[:-1], _flat_named(grads)) <= 1e-5 and rel_l2(flat2[:-1], _flat_named(grads)) <= 1e-5


def test_swae_training_through_the_training_module(golden, tmp_path):
    """training.train with the custom loss: the first epoch (one batch of 48 rows) draws the reference's stream after
    torch.manual_seed(123) and reproduces its loss; the loss goes down over the epochs"""
    from types import SimpleNamespace
    from baler_b200.modules import training
    g = golden("swae.npz")
    cfg = SimpleNamespace(deterministic_algorithm=False, batch_size=48, data_dimension=1, model_type="dense", lr=1e-3, reg_param=0.001,
                          early_stopping=False, early_stopping_patience=100, min_delta=0, lr_scheduler=False, lr_scheduler_patience=50,
                          epochs=20, test_size=0, intermittent_model_saving=False, intermittent_saving_patience=100, RHO=0.05, l1=True,
                          activation_extraction=False, custom_loss_function="loss_function_swae", latent_space_size=10)
    m = models.CFD_dense_AE(64, 10)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sub_sd(g, "sd0").items()})
    torch.manual_seed(123)
    training.train(m, 64, g["x"], g["x"], str(tmp_path), cfg)
    losses = np.load(tmp_path / "loss_data.npy")
    assert abs(losses[0, 0] - float(g["loss"])) <= 1e-5 * float(g["loss"])
    assert np.isfinite(losses).all() and losses[0, -1] < 0.9 * losses[0, 0]
    with pytest.raises(NotImplementedError):
        training.DeviceAdam(models.AE_Dropout_BN(24, 15), 1e-3, 64, swae=True)
