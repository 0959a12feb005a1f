#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (baler v1.4.0) in this container.

TEST INFRASTRUCTURE ONLY.  Runs where `/root/reference` is mounted (the build container);
the GPU box never sees the reference, it only sees the committed fixtures.

    python oracle/gen_golden.py            # rewrites every fixture (deterministic: CPU, fixed seeds)

Two shims are needed to import the reference on this image:
  * `oracle/_shims/matplotlib` - matplotlib is not installed, helper.py:31 imports plotting;
  * `ReduceLROnPlateau(verbose=True)` (utils.py:313-320) is a TypeError on torch 2.11.
Nothing of the reference is copied: its functions are called and their outputs saved.

Fixtures (all float64 unless noted; state-dict tensors stored under "<prefix>/<key>"):
  ae_cms.npz        trained AE(24,15): normalise -> encode -> decode -> renormalise on 256 rows
  ae_train.npz      fresh AE(24,15): loss, grads, Adam after 1 and 3 steps (mse only and mse+l1)
  ae_fit.npz        training.train() for 3 epochs on 4096 rows: loss_data, final weights
  ae_dbn.npz        AE_Dropout_BN(24,15): eval encode/decode; one train step with captured masks
  cli_roundtrip.npz perform_training/compression/decompression in a temp workspace (4096 rows)
  schedules.npz     LRScheduler / EarlyStopping decision sequences
  conv_ae.npz       Conv_AE(5, 250) on 5x5 blocks: eval encode / decode (weights by seed + checksums)
  conv_shapes.npz   Conv_AE on the other two valid block shapes (3x6 with z = 9, 2x8 with z = 4): eval encode / decode
  cfd_dense.npz     CFD_dense_AE(2500, 25) on 50x50 snapshots: encode / decode (weights by seed + checksums)
  conv_train.npz    Conv_AE(5, 250) training on 600 5x5 blocks, batch 300: loss, gradients and parameters after 1 and 3 Adam
                    steps (the four 2000-wide Linear weights as every 101st entry + l2 norm), BatchNorm2d running statistics,
                    then training.train() for 2 epochs: loss_data, final_layer, eval-mode reconstruction of 64 blocks
  swae.npz          CFD_dense_AE(64, 10) (float32), batch 48: utils.loss_function_swae after torch.manual_seed(123) - loss, SWD
                    term, all gradients, parameters after the Adam step; the prior / projection draws replayed from the seed
  eb_deltas.npz     helper.save_error_bounded_requirement on 4 batches of the trained CMS AE + delta re-application
"""
import os
import sys
import shutil
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("BALER_REFERENCE", "/root/reference")
os.environ["CUDA_VISIBLE_DEVICES"] = ""
sys.path[:0] = [os.path.join(HERE, "_shims"), REF, ROOT]

import numpy as np  # noqa: E402
import torch  # noqa: E402

_Plateau = torch.optim.lr_scheduler.ReduceLROnPlateau


class _PlateauNoVerbose(_Plateau):
    def __init__(self, *a, verbose=None, **k):
        super().__init__(*a, **k)


torch.optim.lr_scheduler.ReduceLROnPlateau = _PlateauNoVerbose

from baler import baler as ref_baler  # noqa: E402
from baler.modules import helper as ref_helper  # noqa: E402
from baler.modules import models as ref_models  # noqa: E402
from baler.modules import training as ref_training  # noqa: E402
from baler.modules import utils as ref_utils  # noqa: E402

from baler_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(4)


def sd_np(model, prefix):
    return {f"{prefix}/{k}": v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def grads_np(model, prefix):
    return {f"{prefix}/{k}": p.grad.detach().numpy().copy() for k, p in model.named_parameters()}


def make_config(**over):
    """The shipped CMS config (CMS_project_v1_config.py) with overrides, on the reference's own Config class."""

    class C(ref_helper.Config):
        pass

    c = C
    c.input_path = "unused"
    c.data_dimension = 1
    c.compression_ratio = 1.6
    c.apply_normalization = True
    c.model_name = "AE"
    c.epochs = 3
    c.lr = 0.001
    c.batch_size = 512
    c.early_stopping = True
    c.lr_scheduler = True
    c.save_error_bounded_deltas = False
    c.error_bounded_requirement = 10
    c.early_stopping_patience = 100
    c.min_delta = 0
    c.lr_scheduler_patience = 50
    c.custom_norm = False
    c.reg_param = 0.001
    c.RHO = 0.05
    c.test_size = 0
    c.extra_compression = False
    c.intermittent_model_saving = False
    c.intermittent_saving_patience = 100
    c.mse_avg = False
    c.mse_sum = True
    c.emd = False
    c.l1 = True
    c.activation_extraction = False
    c.deterministic_algorithm = False
    c.convert_to_blocks = False
    c.separate_model_saving = False
    c.latent_space_size = 15
    for k, v in over.items():
        setattr(c, k, v)
    return c


def gen_ae_cms():
    table = synth.cms_table(20000)
    feats = ref_helper.data_processing.find_minmax(table)
    norm = ref_helper.normalize(table, False)
    assert norm.dtype == np.float32
    torch.manual_seed(0)
    model = ref_models.AE(24, 15)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    xs = torch.tensor(norm, dtype=torch.float64)
    model.train()
    for _ in range(2):  # two epochs of the reference's own loss so activations are in a trained regime
        for i in range(0, len(xs), 512):
            opt.zero_grad()
            xb = xs[i : i + 512]
            loss, _, _ = ref_utils.mse_sum_loss_l1(list(model.children()), xb, model(xb), 0.001, True)
            loss.backward()
            opt.step()
    model.eval()
    with torch.no_grad():
        x = xs[:256]
        z = model.encode(x)
        y = model.decode(z)
    un = ref_helper.renormalize(y.numpy(), feats[0], feats[1])
    d = sd_np(model, "sd")
    d.update(x_raw=table[:256], col_min=table.min(0), col_max=table.max(0), norm_features=feats,
             x_norm=norm[:256], latent=z.numpy(), recon=y.numpy(), unnorm=un)
    np.savez(os.path.join(OUT, "ae_cms.npz"), **d)
    print("ae_cms", z.shape, float(np.abs(z).max()), float(np.abs(y).max()))


def gen_ae_train():
    table = synth.cms_table(2048, seed=7)
    norm = ref_helper.normalize(table, False)
    xs = torch.tensor(norm, dtype=torch.float64)
    d = dict(x_norm=norm)
    for tag, validate in (("mse", True), ("l1", False)):
        torch.manual_seed(0)
        model = ref_models.AE(24, 15)
        if tag == "mse":
            d.update(sd_np(model, "sd0"))
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        model.train()
        losses = []
        for step in range(3):
            xb = xs[step * 512 : (step + 1) * 512]
            opt.zero_grad()
            loss, mse, l1 = ref_utils.mse_sum_loss_l1(list(model.children()), xb, model(xb), 0.001, validate)
            loss.backward()
            if step == 0:
                d.update(grads_np(model, f"g_{tag}"))
                d[f"mse0_{tag}"] = float(mse) if not validate else float(loss)
                d[f"l1_0_{tag}"] = float(l1)
            opt.step()
            losses.append(float(loss))
            if step in (0, 2):
                d.update(sd_np(model, f"sd{step + 1}_{tag}"))
        d[f"losses_{tag}"] = np.array(losses)
    # ragged last batch (448 rows in T600k; here 100 rows) - loss only
    torch.manual_seed(0)
    model = ref_models.AE(24, 15)
    xb = xs[:100]
    loss, _, _ = ref_utils.mse_sum_loss_l1(list(model.children()), xb, model(xb), 0.001, True)
    d["loss_ragged100"] = float(loss)
    np.savez(os.path.join(OUT, "ae_train.npz"), **d)
    print("ae_train", d["losses_mse"], d["losses_l1"])


def gen_ae_fit():
    table = synth.cms_table(4096, seed=11)
    norm = ref_helper.normalize(table, False)
    cfg = make_config(epochs=3)
    tmp = tempfile.mkdtemp()
    try:
        torch.manual_seed(0)
        model = ref_models.AE(24, 15)
        d = {}
        trained = ref_training.train(model, 24, norm, norm, tmp, cfg)
        d["loss_data"] = np.load(os.path.join(tmp, "loss_data.npy"))
        d.update(sd_np(trained, "sd_final"))
    finally:
        shutil.rmtree(tmp)
    np.savez(os.path.join(OUT, "ae_fit.npz"), **d)
    print("ae_fit", d["loss_data"])


def gen_ae_dbn():
    table = synth.cms_table(2048, seed=13)
    norm = ref_helper.normalize(table, False)
    xs = torch.tensor(norm, dtype=torch.float64)
    torch.manual_seed(0)
    model = ref_models.AE_Dropout_BN(24, 15)
    # make BN buffers / affine non-trivial so folding is actually tested
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.copy_(0.5 + torch.rand(m.weight.shape, generator=g, dtype=torch.float64))
                m.bias.copy_(0.2 * torch.randn(m.bias.shape, generator=g, dtype=torch.float64))
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g, dtype=torch.float64))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g, dtype=torch.float64))
    d = dict(x_norm=norm[:512])
    d.update(sd_np(model, "sd0"))
    model.eval()
    with torch.no_grad():
        z = model.encode(xs[:256])
        y = model.decode(z)
    d.update(latent_eval=z.numpy(), recon_eval=y.numpy())
    # one train step; dropout masks captured from torch so the other side can inject them
    masks = {}

    def hook(name):
        def f(mod, inp, out):
            masks[name] = (out != 0).numpy()
        return f

    hs = [m.register_forward_hook(hook(f"mask{i}")) for i, m in enumerate(
        [m for m in model.enc_nn if isinstance(m, torch.nn.Dropout)])]
    model.train()
    torch.manual_seed(5)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    xb = xs[:512]
    opt.zero_grad()
    out = model(xb)
    loss, _, _ = ref_utils.mse_sum_loss_l1(list(model.children()), xb, out, 0.001, True)
    loss.backward()
    d.update(grads_np(model, "g"))
    opt.step()
    for h in hs:
        h.remove()
    d.update({k: v for k, v in masks.items()})
    d.update(recon_train=out.detach().numpy(), loss_train=float(loss))
    d.update(sd_np(model, "sd1"))
    np.savez_compressed(os.path.join(OUT, "ae_dbn.npz"), **d)
    print("ae_dbn", float(loss), {k: v.mean() for k, v in masks.items()})


def gen_cli_roundtrip():
    """`baler --mode train|compress|decompress` on a temp copy of the CMS project layout."""
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    try:
        os.chdir(tmp)
        ref_helper.create_new_project("CMS_workspace", "CMS_project_v1")
        table = synth.cms_table(4096, seed=17)
        path = os.path.join("workspaces", "CMS_workspace", "data", "example_CMS_data.npz")
        np.savez(path, data=table, names=synth.CMS_NAMES)
        type_list = ["float64"] * 12 + ["int"] * 7 + ["float64"] * 3 + ["int"] * 2
        cfg = make_config(input_path=path, epochs=2, type_list=type_list, activation_extraction=True)
        if hasattr(cfg, "latent_space_size"):
            del cfg.latent_space_size
        out = os.path.join("workspaces", "CMS_workspace", "CMS_project_v1", "output")
        torch.manual_seed(0)
        ref_baler.perform_training(out, cfg, False)
        ref_baler.perform_compression(out, cfg, False)
        ref_baler.perform_decompression(out, cfg, False)
        sd = torch.load(os.path.join(out, "compressed_output", "model.pt"))
        d = {f"sd/{k}": v.numpy() for k, v in sd.items()}
        comp = np.load(os.path.join(out, "compressed_output", "compressed.npz"))
        dec = np.load(os.path.join(out, "decompressed_output", "decompressed.npz"))
        d.update(compressed=comp["data"], comp_norm_features=comp["normalization_features"],
                 decompressed=dec["data"],
                 norm_features=np.load(os.path.join(out, "training", "normalization_features.npy")),
                 loss_data=np.load(os.path.join(out, "training", "loss_data.npy")),
                 activations=np.load(os.path.join(out, "training", "activations.npy")))
        print("cli", comp["data"].shape, comp["data"].dtype, dec["data"].shape, dec["data"].dtype, d["loss_data"])
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp)
    np.savez(os.path.join(OUT, "cli_roundtrip.npz"), **d)


def gen_schedules():
    """LR plateau schedule (tests/test_utils.py:83-108 pattern) and early stopping decisions."""
    d = {}
    lin = torch.nn.Linear(10, 1)
    opt = torch.optim.SGD(lin.parameters(), lr=0.1)
    sched = ref_utils.LRScheduler(opt, patience=2, min_lr=1e-5, factor=0.5)
    seq = [10.0, 9.0, 8.0, 7.0] + [10.0, 9.0, 10.0, 11.0, 12.0] + [10.0] * 100
    lrs = []
    for v in seq:
        sched(v)
        lrs.append(opt.param_groups[0]["lr"])
    d.update(lr_losses=np.array(seq), lr_values=np.array(lrs))
    # defaults used by training.train: patience from config, factor 0.5, min_lr 1e-6
    opt = torch.optim.Adam(lin.parameters(), lr=1e-3)
    sched = ref_utils.LRScheduler(opt, patience=3)
    rng = np.random.default_rng(3)
    seq2 = list(np.abs(1.0 + 0.05 * rng.normal(size=80)))
    lrs2 = []
    for v in seq2:
        sched(v)
        lrs2.append(opt.param_groups[0]["lr"])
    d.update(lr2_losses=np.array(seq2), lr2_values=np.array(lrs2))
    es = ref_utils.EarlyStopping(patience=4, min_delta=0.01)
    seq3 = [1.0, 0.9, 0.95, 0.895, 0.7, 0.71, 0.72, 0.695, 0.73, 0.74]
    stops, counters = [], []
    for v in seq3:
        es(v)
        stops.append(es.early_stop)
        counters.append(es.counter)
    d.update(es_losses=np.array(seq3), es_stop=np.array(stops), es_counter=np.array(counters))
    np.savez(os.path.join(OUT, "schedules.npz"), **d)
    print("schedules", lrs[-1], lrs2[-1], stops)


def _checksums(model):
    return {f"chk/{k}": np.array(float(v.double().abs().sum())) for k, v in model.state_dict().items()}


def randomise_bn2d(model, seed=2):
    """non-trivial BatchNorm2d buffers / affine so that folding is actually tested; the test applies the same"""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(0.5 + torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(0.2 * torch.randn(m.bias.shape, generator=g))
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))


def gen_conv_shapes():
    """the other two block shapes whose conv stack flattens to the hard-coded 128 values: 3x6 blocks of 48x48
    snapshots with Conv_AE(6, 9), 2x8 blocks with Conv_AE(8, 4); eval encode / decode, weights by seed + checksums"""
    d = {}
    for tag, (h, w), z_dim in (("b36", (3, 6), 9), ("b28", (2, 8), 4)):
        snaps = synth.cfd_snapshots(2, h=48, w=48)  # 2 x 48 x 48 -> 256 blocks of 3x6 / 288 of 2x8
        blocks = ref_helper.data_processing.convert_to_blocks_util([1, h, w], snaps)
        xs = torch.tensor(blocks, dtype=torch.float32).view(blocks.shape[0], 1, h, w)
        torch.manual_seed(0)
        model = ref_models.Conv_AE(w, z_dim)
        randomise_bn2d(model)
        d[f"{tag}/blocks"] = blocks.astype(np.float32)
        d.update({f"{tag}/{k}": v for k, v in _checksums(model).items()})
        model.eval()
        with torch.no_grad():
            z = model.encode(xs)
            y = model.decode(z)
        d[f"{tag}/latent_eval"], d[f"{tag}/recon_eval"] = z.numpy(), y.numpy()
        d[f"{tag}/final_layer"] = np.array(model.get_final_layer_dims())
        print("conv_shapes", tag, z.shape, y.shape)
    np.savez_compressed(os.path.join(OUT, "conv_shapes.npz"), **d)
    print("conv_shapes", os.path.getsize(os.path.join(OUT, "conv_shapes.npz")) / 1e6, "MB")


def gen_conv_ae():
    """Conv_AE(5, 250) on 5x5 blocks (the only block shapes the reference model accepts).  The 6 MB of
    weights are not stored: they are torch.manual_seed(0) initial weights (reproduced by the drop-in class, checked
    through per-tensor checksums) plus the BatchNorm2d statistics written here."""
    snaps = synth.cfd_snapshots(6)  # 6 x 50 x 50 -> 600 blocks of 1x5x5
    blocks = ref_helper.data_processing.convert_to_blocks_util([1, 5, 5], snaps)
    xs = torch.tensor(blocks, dtype=torch.float32).view(blocks.shape[0], 1, 5, 5)
    torch.manual_seed(0)
    model = ref_models.Conv_AE(5, 250)
    randomise_bn2d(model)
    d = dict(blocks=blocks.astype(np.float32))
    d.update(_checksums(model))
    d.update({f"bn/{k}": v.numpy().copy() for k, v in model.state_dict().items()
              if ".3." in k and k.startswith("q_z_conv") or k.startswith(("p_x_conv.1.", "p_x_conv.4."))})
    model.eval()
    with torch.no_grad():
        z = model.encode(xs)
        y = model.decode(z)
    d.update(latent_eval=z.numpy(), recon_eval=y.numpy())
    np.savez_compressed(os.path.join(OUT, "conv_ae.npz"), **d)
    print("conv_ae", z.shape, y.shape, os.path.getsize(os.path.join(OUT, "conv_ae.npz")) / 1e6, "MB")


CONV_BIG = ("q_z_lin.0.weight", "q_z_lin.2.weight", "p_x_lin.0.weight", "p_x_lin.2.weight")
CONV_STRIDE = 101


def _conv_pack(named, prefix):
    """tensors by name; the four big Linear weights as a strided sample + l2 norm (6 MB each step otherwise)"""
    d = {}
    for k, v in named:
        a = v.detach().double().numpy().reshape(-1) if v.dim() else v.detach().numpy().copy()
        if k in CONV_BIG:
            d[f"{prefix}/{k}#sample"] = a[::CONV_STRIDE].copy()
            d[f"{prefix}/{k}#norm"] = np.array(np.sqrt((a * a).sum()))
        else:
            d[f"{prefix}/{k}"] = a.copy() if v.dim() else a
    return d


def conv_train_blocks():
    snaps = synth.cfd_snapshots(6)
    blocks = ref_helper.data_processing.convert_to_blocks_util([1, 5, 5], snaps).astype(np.float32)
    return ((blocks - blocks.min()) / (blocks.max() - blocks.min())).astype(np.float32)  # 600 x 5 x 5 in [0, 1]


def gen_conv_train():
    """Conv_AE training steps exactly as training.fit runs them (training.py:62-95: forward in train mode,
    mse_sum_loss_l1(validate=True) - whose divisor is true_data.shape[1] = 1 for (B, 1, 5, 5) batches -, backward, Adam),
    from torch.manual_seed(0) initial weights + randomise_bn2d (reproduced by the test through the drop-in class)."""
    blocks = conv_train_blocks()
    xs = torch.tensor(blocks, dtype=torch.float32).view(-1, 1, 5, 5)
    d = dict(blocks=blocks.reshape(-1, 25))
    torch.manual_seed(0)
    model = ref_models.Conv_AE(5, 250)
    randomise_bn2d(model)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    model.train()
    losses = []
    for step, lo in enumerate((0, 300, 0)):
        xb = xs[lo:lo + 300]
        opt.zero_grad()
        loss, _, _ = ref_utils.mse_sum_loss_l1(list(model.children()), xb, model(xb), 0.001, True)
        loss.backward()
        if step == 0:
            d.update(_conv_pack([(k, p.grad) for k, p in model.named_parameters()], "g0"))
        opt.step()
        losses.append(float(loss))
        if step in (0, 2):
            d.update(_conv_pack(model.state_dict().items(), f"sd{step + 1}"))
    d["losses"] = np.array(losses)
    model.eval()
    with torch.no_grad():
        d["eval_loss"] = np.array(float(ref_utils.mse_sum_loss_l1(None, xs[:300], model(xs[:300]), 0.001, True)[0]))
    # training.train(): 2 epochs, batch 300, validation on the training blocks (test_size = 0 -> val = train loss)
    cfg = make_config(epochs=2, batch_size=300, data_dimension=2, model_type="convolutional", model_name="Conv_AE",
                      early_stopping=False, lr_scheduler=False, convert_to_blocks=[1, 5, 5], latent_space_size=250)
    tmp = tempfile.mkdtemp()
    try:
        torch.manual_seed(0)
        model = ref_models.Conv_AE(5, 250)
        randomise_bn2d(model)
        trained = ref_training.train(model, 5, blocks, blocks, tmp, cfg)
        d["loss_data"] = np.load(os.path.join(tmp, "loss_data.npy"))
        d["final_layer"] = np.load(os.path.join(tmp, "final_layer.npy"), allow_pickle=True)
        d.update(_conv_pack(trained.state_dict().items(), "sd_final"))
        trained.eval()
        with torch.no_grad():
            d["recon_final"] = trained(xs[:64]).numpy().reshape(64, 25)
    finally:
        shutil.rmtree(tmp)
    np.savez_compressed(os.path.join(OUT, "conv_train.npz"), **d)
    print("conv_train", d["losses"], d["eval_loss"], d["loss_data"], d["final_layer"],
          os.path.getsize(os.path.join(OUT, "conv_train.npz")) / 1e6, "MB")


def gen_swae():
    """One training step with config.custom_loss_function = "loss_function_swae" exactly as training.fit runs it
    (training.py:62-92): reconstructions = model(inputs); z = model.encode(inputs); utils.loss_function_swae(inputs, z,
    reconstructions, latent_dim).  Its random inputs come from torch's global generator in the order randn_like(z), then
    randn(2000, latent_dim) (utils.py:56-90): seeded here and replayed into the fixture."""
    rng = np.random.default_rng(5)
    x = rng.random((48, 64), dtype=np.float32)
    torch.manual_seed(0)
    model = ref_models.CFD_dense_AE(64, 10)
    d = dict(x=x)
    d.update(sd_np(model, "sd0"))
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    model.train()
    xb = torch.from_numpy(x)
    opt.zero_grad()
    recon = model(xb)
    z = model.encode(xb)
    torch.manual_seed(123)
    loss, mse, swd = ref_utils.loss_function_swae(xb, z, recon, 10)
    loss.backward()
    d.update(grads_np(model, "g"))
    opt.step()
    d.update(sd_np(model, "sd1"))
    d.update(loss=float(loss), mse=float(mse), swd=float(swd), latent=z.detach().numpy())
    torch.manual_seed(123)
    d["prior"] = torch.randn_like(z).numpy()
    d["proj"] = ref_utils.get_random_projections("normal", 10, 2000).numpy()
    np.savez_compressed(os.path.join(OUT, "swae.npz"), **d)
    print("swae", d["loss"], d["mse"], d["swd"], os.path.getsize(os.path.join(OUT, "swae.npz")) / 1e6, "MB")
    try:  # the float64 AE: z (double) @ projections (float) - does upstream run at all?
        m64 = ref_models.AE(24, 15)
        x64 = torch.rand(8, 24, dtype=torch.float64)
        ref_utils.loss_function_swae(x64, m64.encode(x64), m64(x64), 15)
        print("swae on the float64 AE: runs")
    except Exception as e:
        print("swae on the float64 AE upstream:", type(e).__name__, str(e)[:100])


def gen_cfd_dense():
    """CFD_dense_AE(2500, 25) (float32) on 50x50 snapshots, as CFD_project_animation configures it."""
    snaps = synth.cfd_snapshots(60)
    xs = torch.tensor(snaps, dtype=torch.float32).view(60, 2500)
    torch.manual_seed(0)
    model = ref_models.CFD_dense_AE(2500, 25)
    d = dict(_checksums(model))
    model.eval()
    with torch.no_grad():
        z = model.encode(xs)
        y = model.decode(z)
    d.update(latent=z.numpy(), recon=y.numpy())
    np.savez_compressed(os.path.join(OUT, "cfd_dense.npz"), **d)
    print("cfd_dense", z.shape, y.shape, os.path.getsize(os.path.join(OUT, "cfd_dense.npz")) / 1e6, "MB")


def gen_eb_deltas():
    """helper.save_error_bounded_requirement (helper.py:442-470) on the trained CMS AE: per batch of normalised rows,
    which elements exceed the relative-error bound and the float16 deltas stored for them (2 batches of 256 rows); then the reference's own
    delta re-application of helper.decompress (helper.py:708-718) on those batches."""
    g = np.load(os.path.join(OUT, "ae_cms.npz"))
    model = ref_models.AE(24, 15)
    model.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")})
    model.eval()
    table = synth.cms_table(512, seed=41)
    norm = ref_helper.normalize(table, False)
    cfg = make_config()
    cfg.error_bounded_requirement = 25.0  # percent
    d = dict(table=table, bound=np.float64(cfg.error_bounded_requirement))
    with torch.no_grad():
        for b in range(2):
            data = torch.tensor(norm[256 * b:256 * (b + 1)], dtype=torch.float64)
            dec = model.decode(model.encode(data)).numpy()
            deltas, index = ref_helper.save_error_bounded_requirement(cfg, dec, data.numpy())
            rows, cols = index
            fixed = dec.copy()
            for i in range(len(rows)):  # helper.py:714-716
                fixed[rows[i]][cols[i]] -= deltas[i]
            d["rows%d" % b], d["cols%d" % b] = rows.astype(np.int64), cols.astype(np.int64)
            d["deltas%d" % b] = np.array(deltas, dtype=np.float16)
            d["decoded%d" % b], d["fixed%d" % b] = dec, fixed
    np.savez_compressed(os.path.join(OUT, "eb_deltas.npz"), **d)
    print("eb_deltas", [len(d["rows%d" % b]) for b in range(2)])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["ae_cms", "ae_train", "ae_fit", "ae_dbn", "cli_roundtrip", "schedules", "conv_ae", "conv_shapes", "conv_train", "cfd_dense", "swae", "eb_deltas"]
    for name in which:
        globals()["gen_" + name]()
