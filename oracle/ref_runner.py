"""Run the staged, unmodified reference (oracle/_ref/baler, see oracle/stage_ref.py) on the host CPU.

TEST INFRASTRUCTURE - NOT PRODUCT CODE: only bench.py's CPU legs (`--impl reference`, `cpu_baseline`) import this.
The reference picks cuda:0 when it sees one (helper.py:425-439): callers hide the GPUs (CUDA_VISIBLE_DEVICES="") BEFORE
torch is imported, which is why the bench runs these legs in a subprocess.

Timed items follow BASELINE.md section 3:
  C1 helper.compress(model_path, config) as shipped          (helper.py:473-616)
  C2 helper.decompress(...) as shipped + helper.renormalize   (helper.py:619-733, baler.py:410-424)
  C3 bare model.encode / model.decode over the whole tensor in one call (float64 as shipped, and .float())
  C4 one pass of training.fit (training.py:31-101) for AE and AE_Dropout_BN, bs 512, Adam lr 1e-3
  C5 Conv_AE encode / decode on 5 x 5 blocks, batch 600       (models.py:316-407)
"""
import contextlib
import io
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_mods = None


def available():
    return os.path.isdir(os.path.join(REF_DIR, "baler"))


def load():
    """import the staged reference with the two run-time shims; returns a namespace of its modules"""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("oracle/_ref/baler is not staged (python oracle/stage_ref.py in the build container)")
    sys.path[:0] = [os.path.join(HERE, "_shims"), REF_DIR]
    import torch

    plateau = torch.optim.lr_scheduler.ReduceLROnPlateau
    if not getattr(plateau, "_baler_shim", False):
        class PlateauNoVerbose(plateau):
            _baler_shim = True

            def __init__(self, *a, verbose=None, **k):
                super().__init__(*a, **k)

        torch.optim.lr_scheduler.ReduceLROnPlateau = PlateauNoVerbose
    from baler import baler as b
    from baler.modules import helper, models, training, utils

    class NS:
        pass

    ns = NS()
    ns.baler, ns.helper, ns.models, ns.training, ns.utils, ns.torch = b, helper, models, training, utils, torch
    assert os.path.realpath(helper.__file__).startswith(os.path.realpath(REF_DIR)), helper.__file__
    _mods = ns
    return ns


def make_config(ns, **over):
    """the shipped CMS_project_v1 config on the reference's own Config class"""

    class C(ns.helper.Config):
        pass

    c = C
    base = dict(input_path="unused", data_dimension=1, compression_ratio=1.6, apply_normalization=True, model_name="AE",
                model_type="dense", epochs=1, lr=0.001, batch_size=512, early_stopping=True, lr_scheduler=True,
                save_error_bounded_deltas=False, error_bounded_requirement=10, early_stopping_patience=100, min_delta=0,
                lr_scheduler_patience=50, custom_norm=False, reg_param=0.001, RHO=0.05, test_size=0,
                extra_compression=False, intermittent_model_saving=False, intermittent_saving_patience=100, mse_avg=False,
                mse_sum=True, emd=False, l1=True, activation_extraction=False, deterministic_algorithm=False,
                convert_to_blocks=False, separate_model_saving=False)
    base.update(over)
    for k, v in base.items():
        setattr(c, k, v)
    return c


@contextlib.contextmanager
def _quiet():
    """the reference prints progress (stdout) and tqdm bars (stderr); bench.py's stdout carries exactly one JSON line"""
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield


def compress_decompress_as_shipped(table, names, state_dict):
    """C1 + C2 on `table` (float32 [n, 24]) with the AE weights `state_dict` (numpy): seconds of helper.compress and of
    helper.decompress + helper.renormalize, and the reconstruction (for a sanity check against the input)"""
    import numpy as np
    ns = load()
    torch = ns.torch
    tmp = tempfile.mkdtemp()
    try:
        inp = os.path.join(tmp, "table.npz")
        np.savez(inp, data=table, names=names)
        model_path = os.path.join(tmp, "model.pt")
        torch.save({k: torch.from_numpy(np.asarray(v)) for k, v in state_dict.items()}, model_path)
        cfg = make_config(ns, input_path=inp)
        with _quiet():
            t0 = time.perf_counter()
            compressed, _, _, _ = ns.helper.compress(model_path, cfg)
            t_c = time.perf_counter() - t0
        feats = ns.helper.data_processing.find_minmax(table)
        comp_path = os.path.join(tmp, "compressed.npz")
        np.savez(comp_path, data=compressed, names=names, normalization_features=feats)
        with _quiet():
            t0 = time.perf_counter()
            dec, _, nf = ns.helper.decompress(model_path, comp_path, None, None, "AE", cfg, tmp, table.shape)
            dec = ns.helper.renormalize(dec, nf[0], nf[1])
            t_d = time.perf_counter() - t0
        return t_c, t_d, np.asarray(dec)
    finally:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)


def bare_encode_decode(table, state_dict, dtype="float64", repeats=2):
    """C3: normalise (vectorised numpy), model.encode and model.decode over the whole tensor in ONE call each: the
    math-only best case of the reference's CPU path; returns seconds (best of `repeats`)"""
    import numpy as np
    ns = load()
    torch = ns.torch
    model = ns.models.AE(table.shape[1], state_dict["en4.weight"].shape[0])
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state_dict.items()})
    model.eval()
    if dtype == "float32":
        model = model.float()
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        mn = table.min(axis=0)
        rg = table.max(axis=0) - mn
        x = torch.from_numpy(((table - mn) / rg).astype(dtype))
        with torch.no_grad():
            z = model.encode(x)
            y = model.decode(z)
        out = y.numpy() * rg + mn
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, out


def fit_pass(model_name, x_norm, batch_size=512, lr=1e-3):
    """C4: one call of training.fit (the reference's epoch loop, DataLoader as training.train builds it) on the normalised
    rows `x_norm`; returns (seconds, epoch loss)"""
    import numpy as np
    from torch.utils.data import DataLoader
    ns = load()
    torch = ns.torch
    torch.manual_seed(0)
    n_features = x_norm.shape[1]
    z = 15 if n_features == 24 else max(1, n_features // 2)
    model = getattr(ns.models, model_name)(n_features, z)
    cfg = make_config(ns, model_name=model_name, latent_space_size=z)
    ds = torch.tensor(np.asarray(x_norm), dtype=torch.float64)
    dl = DataLoader(ds, batch_size=batch_size, shuffle=False, drop_last=False)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    with _quiet():
        t0 = time.perf_counter()
        out = ns.training.fit(config=cfg, model=model, train_dl=dl, model_children=list(model.children()),
                              regular_param=cfg.reg_param, optimizer=opt, latent_dim=z, RHO=cfg.RHO, l1=cfg.l1,
                              n_dimensions=cfg.data_dimension)
        dt = time.perf_counter() - t0
    return dt, float(out[0])


def conv_encode_decode(blocks, z_dim=250, batch=600):
    """C5: Conv_AE(n_features, z_dim) eval-mode encode / decode of [n, 1, 5, 5] float32 blocks in batches of `batch`;
    returns (encode seconds, decode seconds)"""
    ns = load()
    torch = ns.torch
    torch.manual_seed(0)
    model = ns.models.Conv_AE(5, z_dim).eval()
    x = torch.from_numpy(blocks)
    n = (x.shape[0] // batch) * batch
    with torch.no_grad():
        model.encode(x[:batch])
        t0 = time.perf_counter()
        zs = [model.encode(x[i:i + batch]) for i in range(0, n, batch)]
        t_e = time.perf_counter() - t0
        t0 = time.perf_counter()
        for z in zs:
            model.decode(z)
        t_d = time.perf_counter() - t0
    return t_e, t_d, n
