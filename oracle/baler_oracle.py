"""CPU restatement (numpy, float64) of Baler's autoencoder train / compress / decompress path.

TEST INFRASTRUCTURE - NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module.  The product
path (`baler_b200/`) never does and fails loudly if its CUDA library is missing.

Parity status: PINNED.  Every function here is checked in `tests/test_oracle_golden.py`
against outputs of the reference itself (baler v1.4.0 run in the build container by
`oracle/gen_golden.py`, fixtures in `tests/golden/`) and against the known-answer vectors of
the reference's own unit tests (`tests/test_data_processing.py:52-121`, `tests/test_utils.py:83-108`).

All arithmetic of this path is executed upstream by PyTorch (pinned 2.2.1 in the reference's
`poetry.lock`), which is not part of `/root/reference`; the published semantics of
nn.Linear, F.leaky_relu, nn.Dropout, nn.BatchNorm1d, nn.MSELoss(reduction="sum"),
torch.optim.Adam and lr_scheduler.ReduceLROnPlateau are restated below.  Citations
`file:line` are relative to `/root/reference/`.
"""
import math

import numpy as np

LEAKY_SLOPE = 0.01  # F.leaky_relu default, baler/modules/models.py:142
BN_EPS = 1e-5  # nn.BatchNorm1d default
BN_MOMENTUM = 0.1
DROPOUT_P = (0.5, 0.4, 0.3, 0.2)  # baler/modules/models.py:263-275

AE_LAYERS = ("en1", "en2", "en3", "en4", "de1", "de2", "de3", "de4")  # models.py:128-136
HIDDEN = (200, 100, 50)


# --------------------------------------------------------------------------- normalisation
def find_minmax(data):
    """[min; max-min] per column (axis 0).  baler/modules/data_processing.py:113-130."""
    data = np.asarray(data)
    mx = data.max(axis=0)
    mn = data.min(axis=0)
    return np.array([mn, mx - mn])


def normalize(data, custom_norm=False):
    """(x - min) / (max - min) per column, in the input dtype, true division, no guard for
    range == 0.  baler/modules/helper.py:261-274 applying data_processing.py:133-153 on axis 0."""
    data = np.asarray(data)
    if custom_norm:
        return data
    mn = data.min(axis=0)
    rng = data.max(axis=0) - mn
    return (data - mn) / rng


def renormalize(norm_data, min_list, range_list):
    """y * range + min.  baler/modules/data_processing.py:188-203."""
    return np.asarray(norm_data) * np.asarray(range_list) + np.asarray(min_list)


# --------------------------------------------------------------------------- dense AE
def leaky_relu(x):
    return np.where(x > 0, x, LEAKY_SLOPE * x)


def linear(x, w, b):
    """nn.Linear: x @ W.T + b with W of shape (out, in)."""
    return x @ w.T + b


def init_ae_shapes(n_features, z_dim):
    dims = (n_features,) + HIDDEN + (z_dim,) + HIDDEN[::-1] + (n_features,)
    return {name: (dims[i + 1], dims[i]) for i, name in enumerate(AE_LAYERS)}


def ae_encode(sd, x):
    """baler/modules/models.py:141-145 (no activation on the latent)."""
    h = np.asarray(x, dtype=np.float64)
    for name in AE_LAYERS[:3]:
        h = leaky_relu(linear(h, sd[name + ".weight"], sd[name + ".bias"]))
    return linear(h, sd["en4.weight"], sd["en4.bias"])


def ae_decode(sd, z):
    """baler/modules/models.py:147-152 (no activation on the output)."""
    h = np.asarray(z, dtype=np.float64)
    for name in AE_LAYERS[4:7]:
        h = leaky_relu(linear(h, sd[name + ".weight"], sd[name + ".bias"]))
    return linear(h, sd["de4.weight"], sd["de4.bias"])


def ae_forward(sd, x):
    return ae_decode(sd, ae_encode(sd, x))


def mse_sum_loss(recon, x):
    """sum((recon - x)^2) / n_columns.  baler/modules/utils.py:195-199."""
    return float(np.sum((recon - x) ** 2) / x.shape[1])


def swd_term(z, prior, proj, reg_weight=100.0):
    """utils.compute_swd as utils.loss_function_swae calls it (utils.py:27-76) with p = 2: `prior` = the torch.randn_like(z)
    draws, `proj` [S, D] = utils.get_random_projections (unit rows).  reg_weight / (B (B - 1)) * mean over projections and
    ranks of (sort(z P^T) - sort(prior P^T))^2, sorted along the batch.  Returns (value, d value / d z)."""
    z, prior, proj = (np.asarray(a, dtype=np.float64) for a in (z, prior, proj))
    b, s = z.shape[0], proj.shape[0]
    lat, pri = z @ proj.T, prior @ proj.T                        # [B, S]
    order = np.argsort(lat, axis=0, kind="stable")
    w = np.take_along_axis(lat, order, axis=0) - np.sort(pri, axis=0)
    coef = reg_weight / (b * (b - 1)) / (s * b)
    dlat = np.zeros_like(lat)
    np.put_along_axis(dlat, order, 2.0 * w * coef, axis=0)
    return float(coef * (w * w).sum()), dlat @ proj


def ae_loss_and_grads(sd, x, reg_param=0.0, l1=False, swae=None):
    """Loss and parameter gradients of one `fit` step for the dense AE.

    loss = mse_sum_loss  (+ reg_param * l1_chain when `l1`, i.e. utils.mse_sum_loss_l1 with
    validate=False, utils.py:201-209: a SECOND chain v = relu(child(v)) over the 8 Linears,
    ReLU not LeakyReLU, also after the latent and the output; l1 += mean|v| per layer).
    `training.fit` always passes validate=True (training.py:83-89), so l1=False is what ships.
    swae = (prior, proj[, reg_weight]): config.custom_loss_function == "loss_function_swae" (training.py:70-78): the loss is
    sum-MSE / n_columns + swd_term(latent), and the encoder receives both gradients (upstream runs the encoder a second time
    for z, the same values for this model).
    Returns (loss, mse, l1_value, grads dict keyed like the state dict).
    """
    x = np.asarray(x, dtype=np.float64)
    n_cols = x.shape[1]
    acts = [x]
    pre = []
    h = x
    for i, name in enumerate(AE_LAYERS):
        a = linear(h, sd[name + ".weight"], sd[name + ".bias"])
        pre.append(a)
        h = a if i in (3, 7) else leaky_relu(a)
        acts.append(h)
    recon = h
    mse = float(np.sum((recon - x) ** 2) / n_cols)
    grads = {}
    d = 2.0 * (recon - x) / n_cols
    swd = 0.0
    if swae is not None:
        swd, dz_swd = swd_term(acts[4], *swae)
    for i in range(7, -1, -1):
        name = AE_LAYERS[i]
        if i == 3 and swae is not None:
            d = d + dz_swd
        if i not in (3, 7):
            d = d * np.where(pre[i] > 0, 1.0, LEAKY_SLOPE)
        grads[name + ".weight"] = d.T @ acts[i]
        grads[name + ".bias"] = d.sum(axis=0)
        d = d @ sd[name + ".weight"]
    l1_val = 0.0
    if l1:
        vals = [x]
        v = x
        for name in AE_LAYERS:
            v = np.maximum(linear(v, sd[name + ".weight"], sd[name + ".bias"]), 0.0)
            vals.append(v)
            l1_val += float(np.mean(np.abs(v)))
        d = np.zeros_like(vals[-1])
        for i in range(7, -1, -1):
            name = AE_LAYERS[i]
            v = vals[i + 1]
            d = (d + reg_param / v.size) * (v > 0)  # d mean|v| / dv = 1/size where v > 0
            grads[name + ".weight"] = grads[name + ".weight"] + d.T @ vals[i]
            grads[name + ".bias"] = grads[name + ".bias"] + d.sum(axis=0)
            d = d @ sd[name + ".weight"]
    loss = mse + (reg_param * l1_val if l1 else 0.0) + swd
    return loss, mse, l1_val, grads


# --------------------------------------------------------------------------- AE_Dropout_BN
DBN_ENC = ("enc_nn.0", "enc_nn.3", "enc_nn.6", "enc_nn.9")  # models.py:261-278
DBN_DEC = ("dec_nn.0", "dec_nn.3", "dec_nn.6", "dec_nn.9")  # models.py:281-298
DBN_BN = ("dec_nn.2", "dec_nn.5", "dec_nn.8", "dec_nn.10")


def dbn_encode(sd, x, masks=None):
    """4 x (Linear -> Dropout -> LeakyReLU), activation also on the latent (models.py:261-278).
    masks: None for eval; else 4 boolean keep-masks (train: out = in * mask / (1 - p))."""
    h = np.asarray(x, dtype=np.float64)
    for i, name in enumerate(DBN_ENC):
        a = linear(h, sd[name + ".weight"], sd[name + ".bias"])
        if masks is not None:
            a = a * masks[i] / (1.0 - DROPOUT_P[i])
        h = leaky_relu(a)
    return h


def _bn_eval(sd, name, x):
    inv = 1.0 / np.sqrt(sd[name + ".running_var"] + BN_EPS)
    return (x - sd[name + ".running_mean"]) * inv * sd[name + ".weight"] + sd[name + ".bias"]


def dbn_decode(sd, z):
    """eval mode: 3 x (Linear -> LeakyReLU -> BN) + Linear -> BN -> ReLU (models.py:281-298)."""
    h = np.asarray(z, dtype=np.float64)
    for i in range(3):
        h = leaky_relu(linear(h, sd[DBN_DEC[i] + ".weight"], sd[DBN_DEC[i] + ".bias"]))
        h = _bn_eval(sd, DBN_BN[i], h)
    h = _bn_eval(sd, DBN_BN[3], linear(h, sd[DBN_DEC[3] + ".weight"], sd[DBN_DEC[3] + ".bias"]))
    return np.maximum(h, 0.0)


def dbn_train_step(sd, x, masks):
    """One train-mode forward + backward of AE_Dropout_BN with injected dropout keep-masks.

    BatchNorm1d in train mode: biased batch variance for the output, running stats updated with
    momentum 0.1 and the UNBIASED variance, num_batches_tracked += 1.
    Returns (loss, recon, grads, new_buffers)."""
    x = np.asarray(x, dtype=np.float64)
    n, n_cols = x.shape
    cache = {}
    h = x
    for i, name in enumerate(DBN_ENC):
        cache["ein%d" % i] = h
        a = linear(h, sd[name + ".weight"], sd[name + ".bias"])
        a = a * masks[i] / (1.0 - DROPOUT_P[i])
        cache["epre%d" % i] = a
        h = leaky_relu(a)
    new_buf = {}
    for i in range(4):
        cache["din%d" % i] = h
        a = linear(h, sd[DBN_DEC[i] + ".weight"], sd[DBN_DEC[i] + ".bias"])
        cache["dpre%d" % i] = a
        u = leaky_relu(a) if i < 3 else a
        mean = u.mean(axis=0)
        var = u.var(axis=0)
        inv = 1.0 / np.sqrt(var + BN_EPS)
        xhat = (u - mean) * inv
        cache["xhat%d" % i], cache["inv%d" % i] = xhat, inv
        bn = DBN_BN[i]
        h = xhat * sd[bn + ".weight"] + sd[bn + ".bias"]
        new_buf[bn + ".running_mean"] = (1 - BN_MOMENTUM) * sd[bn + ".running_mean"] + BN_MOMENTUM * mean
        new_buf[bn + ".running_var"] = (1 - BN_MOMENTUM) * sd[bn + ".running_var"] + BN_MOMENTUM * var * n / (n - 1)
        new_buf[bn + ".num_batches_tracked"] = sd[bn + ".num_batches_tracked"] + 1
        if i == 3:
            cache["bnout3"] = h
            h = np.maximum(h, 0.0)
    recon = h
    loss = float(np.sum((recon - x) ** 2) / n_cols)
    g = {}
    d = 2.0 * (recon - x) / n_cols
    for i in range(3, -1, -1):
        bn = DBN_BN[i]
        if i == 3:
            d = d * (cache["bnout3"] > 0)
        xhat, inv = cache["xhat%d" % i], cache["inv%d" % i]
        g[bn + ".weight"] = (d * xhat).sum(axis=0)
        g[bn + ".bias"] = d.sum(axis=0)
        dxh = d * sd[bn + ".weight"]
        d = inv / n * (n * dxh - dxh.sum(axis=0) - xhat * (dxh * xhat).sum(axis=0))
        if i < 3:
            d = d * np.where(cache["dpre%d" % i] > 0, 1.0, LEAKY_SLOPE)
        g[DBN_DEC[i] + ".weight"] = d.T @ cache["din%d" % i]
        g[DBN_DEC[i] + ".bias"] = d.sum(axis=0)
        d = d @ sd[DBN_DEC[i] + ".weight"]
    for i in range(3, -1, -1):
        d = d * np.where(cache["epre%d" % i] > 0, 1.0, LEAKY_SLOPE)
        d = d * masks[i] / (1.0 - DROPOUT_P[i])
        g[DBN_ENC[i] + ".weight"] = d.T @ cache["ein%d" % i]
        g[DBN_ENC[i] + ".bias"] = d.sum(axis=0)
        d = d @ sd[DBN_ENC[i] + ".weight"]
    return loss, recon, g, new_buf


# --------------------------------------------------------------------------- optimiser / schedules
class Adam:
    """torch.optim.Adam defaults as used at baler/modules/training.py:266
    (betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad); single-tensor update:
    m = lerp(m, g, 1-b1); v = b2 v + (1-b2) g^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.params = params
        self.lr, self.betas, self.eps = lr, betas, eps
        self.m = {k: np.zeros_like(v) for k, v in params.items()}
        self.v = {k: np.zeros_like(v) for k, v in params.items()}
        self.t = 0

    def step(self, grads):
        self.t += 1
        b1, b2 = self.betas
        bc1 = 1.0 - b1 ** self.t
        bc2 = 1.0 - b2 ** self.t
        for k, g in grads.items():
            self.m[k] = self.m[k] + (g - self.m[k]) * (1.0 - b1)
            self.v[k] = b2 * self.v[k] + (1.0 - b2) * g * g
            denom = np.sqrt(self.v[k]) / math.sqrt(bc2) + self.eps
            self.params[k] = self.params[k] - (self.lr / bc1) * self.m[k] / denom


class ReduceLROnPlateau:
    """mode="min", threshold 1e-4 relative, cooldown 0 - what utils.LRScheduler builds
    (baler/modules/utils.py:306-323; defaults factor 0.5, min_lr 1e-6 at utils.py:23-24)."""

    def __init__(self, lr, patience, factor=0.5, min_lr=1e-6, threshold=1e-4, eps=1e-8):
        self.lr, self.patience, self.factor, self.min_lr = lr, patience, factor, min_lr
        self.threshold, self.eps = threshold, eps
        self.best = math.inf
        self.num_bad = 0

    def step(self, metric):
        if metric < self.best * (1.0 - self.threshold):
            self.best = metric
            self.num_bad = 0
        else:
            self.num_bad += 1
        if self.num_bad > self.patience:
            new_lr = max(self.lr * self.factor, self.min_lr)
            if self.lr - new_lr > self.eps:
                self.lr = new_lr
            self.num_bad = 0
        return self.lr


class EarlyStopping:
    """baler/modules/utils.py:248-282 (note: equality with min_delta changes nothing)."""

    def __init__(self, patience, min_delta):
        self.patience, self.min_delta = patience, min_delta
        self.counter, self.best_loss, self.early_stop = 0, None, False

    def __call__(self, loss):
        if self.best_loss is None:
            self.best_loss = loss
        elif self.best_loss - loss > self.min_delta:
            self.best_loss = loss
            self.counter = 0
        elif self.best_loss - loss < self.min_delta:
            self.counter += 1
            if self.counter >= self.patience:
                self.early_stop = True


# --------------------------------------------------------------------------- loops
def fit_epoch(sd, opt, data, batch_size, reg_param=0.0, l1=False):
    """baler/modules/training.py:31-101: sequential batches (shuffle=False, drop_last=False),
    epoch loss = mean of per-batch losses."""
    total, nb = 0.0, 0
    for i in range(0, len(data), batch_size):
        loss, _, _, grads = ae_loss_and_grads(opt.params, data[i : i + batch_size], reg_param, l1)
        opt.step(grads)
        total += loss
        nb += 1
    return total / nb


def train(sd, data, batch_size, epochs, lr=1e-3, lr_patience=None, es_patience=None, min_delta=0):
    """baler/modules/training.py:150-348 with test_size=0 (val loss = train loss)."""
    opt = Adam({k: np.array(v, dtype=np.float64) for k, v in sd.items()}, lr=lr)
    sched = ReduceLROnPlateau(lr, lr_patience) if lr_patience is not None else None
    es = EarlyStopping(es_patience, min_delta) if es_patience is not None else None
    losses = []
    for _ in range(epochs):
        loss = fit_epoch(None, opt, data, batch_size)
        losses.append(loss)
        if sched:
            opt.lr = sched.step(loss)
        if es:
            es(loss)
            if es.early_stop:
                break
    return opt.params, np.array([losses, losses])


def compress(sd, table, batch_size=None, as_shipped=False):
    """baler/modules/helper.py:473-616 for the dense AE: normalise with THIS table's min/max,
    upcast to float64, encode.  as_shipped=True keeps the per-batch loop with np.concatenate
    growth (helper.py:584-611)."""
    norm = normalize(table)
    x = norm.astype(np.float64)
    if not as_shipped:
        return ae_encode(sd, x)
    out = None
    for i in range(0, len(x), batch_size):
        z = ae_encode(sd, x[i : i + batch_size])
        out = z if out is None else np.concatenate((out, z))
    return out


def decompress(sd, latent, norm_features=None, type_list=None, batch_size=None, as_shipped=False):
    """baler/modules/helper.py:619-733 + baler/baler.py:410-435: decode, un-normalise with the
    TRAINING features, then the per-column type cast (astype("int") truncates toward zero, result
    stored back into the float64 array)."""
    if not as_shipped:
        out = ae_decode(sd, latent)
    else:
        out = None
        for i in range(0, len(latent), batch_size):
            y = ae_decode(sd, latent[i : i + batch_size])
            out = y if out is None else np.concatenate((out, y))
    if norm_features is not None:
        out = renormalize(out, norm_features[0], norm_features[1])
    if type_list is not None:
        out = np.array(out, dtype=np.float64)
        for c, t in enumerate(type_list):
            out[:, c] = out[:, c].astype(t)
    return out


# --------------------------------------------------------------------------- error-bounded deltas
def error_bounded_deltas(decoded, data, bound_percent):
    """helper.save_error_bounded_requirement (helper.py:442-470) on one batch of NORMALISED rows.

    relative error in percent (decoded - data) / data * 100; +-inf (data == 0) is set to 0, NaN stays NaN and never
    exceeds; the stored delta is np.subtract(decoded, data, dtype=float16), i.e. BOTH operands are rounded to float16
    first and subtracted in float16.  Returns (rows, cols, deltas float16) in np.where (row-major) order.  (Upstream
    leaves `deltas` unbound when nothing exceeds the bound - a crash this restatement does not reproduce.)"""
    decoded = np.asarray(decoded, dtype=np.float64)
    data = np.asarray(data, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        err = (decoded - data) / data * 100.0
    err[np.isinf(err)] = 0.0
    rows, cols = np.where(np.abs(err) > bound_percent)
    deltas = (decoded.astype(np.float16) - data.astype(np.float16))[rows, cols].astype(np.float16)
    return rows, cols, deltas


def apply_error_bounded_deltas(decoded, rows, cols, deltas):
    """helper.decompress (helper.py:708-718): out[row][col] -= delta for every stored delta of the batch"""
    out = np.array(decoded, copy=True)
    for r, c, d in zip(rows, cols, deltas):
        out[r][c] -= d
    return out


# --------------------------------------------------------------------------- Conv_AE training
def conv_chain_train_step(spec, x, loss_columns=1):
    """One train-mode forward + loss + backward of models.Conv_AE (models.py:316-407) on flattened blocks x [B, H*W], as
    training.fit runs it (training.py:62-92).  `spec` describes the network as a chain of affine layers (what the product's
    Conv_AE.training_spec returns, plain numpy): dims, acts ("relu" / "none"), weights[l] / biases[l] (a Linear's (out, in)
    matrix, or a (transposed) convolution's flat kernel with w_maps[l][i] = kernel index of dense entry i, -1 = zero, and a
    per-channel bias), bn[l] = None or (4, C) gamma / beta / running_mean / running_var of the nn.BatchNorm2d that follows.
    nn.BatchNorm2d in train mode: per-channel mean and BIASED variance over batch x positions, eps 1e-5.
    Loss: nn.MSELoss(reduction="sum") / true_data.shape[1] (utils.py:195-199), and shape[1] of a (B, 1, H, W) batch is 1.
    Returns (loss, grads) with grads[l] = dict(weight, bias[, gamma, beta]) in the layers' TRAINABLE parametrisation, plus
    the batch statistics [(mean, biased var)] per BatchNorm layer."""
    x = np.asarray(x, dtype=np.float64)
    L = len(spec["weights"])
    dense, a, saved, stats = [], x, [], []
    for l in range(L):
        k, n = spec["dims"][l], spec["dims"][l + 1]
        w = np.asarray(spec["weights"][l], dtype=np.float64)
        b = np.asarray(spec["biases"][l], dtype=np.float64)
        mp = spec["w_maps"][l]
        wd = w.reshape(n, k) if mp is None else np.where(mp >= 0, w.reshape(-1)[np.clip(mp, 0, None)], 0.0).reshape(n, k)
        bd = np.repeat(b, n // b.size)
        dense.append(wd)
        z = a @ wd.T + bd
        bn = spec["bn"][l]
        xhat = rstd = None
        if bn is not None:
            c = bn.shape[1]
            zc = z.reshape(z.shape[0], c, n // c)
            mean, var = zc.mean(axis=(0, 2)), zc.var(axis=(0, 2))
            stats.append((mean, var))
            rstd = 1.0 / np.sqrt(var + BN_EPS)
            xhat = (zc - mean[None, :, None]) * rstd[None, :, None]
            z = (xhat * bn[0][None, :, None] + bn[1][None, :, None]).reshape(z.shape)
        out = np.maximum(z, 0.0) if spec["acts"][l] == "relu" else z
        saved.append((a, out, xhat, rstd))
        a = out
    diff = a - x
    loss = float((diff * diff).sum() / loss_columns)
    d = 2.0 * diff / loss_columns
    grads = [None] * L
    for l in reversed(range(L)):
        a_in, out, xhat, rstd = saved[l]
        n = spec["dims"][l + 1]
        if spec["acts"][l] == "relu":
            d = d * (out > 0)
        g = {}
        bn = spec["bn"][l]
        if bn is not None:
            c = bn.shape[1]
            dc = d.reshape(d.shape[0], c, n // c)
            g["gamma"], g["beta"] = (dc * xhat).sum(axis=(0, 2)), dc.sum(axis=(0, 2))
            m = dc.shape[0] * dc.shape[2]
            dc = (bn[0] * rstd)[None, :, None] * (dc - g["beta"][None, :, None] / m - xhat * g["gamma"][None, :, None] / m)
            d = dc.reshape(d.shape)
        gw, gb = d.T @ a_in, d.sum(axis=0)
        mp = spec["w_maps"][l]
        nb = np.asarray(spec["biases"][l]).size
        g["bias"] = gb.reshape(nb, n // nb).sum(axis=1)
        if mp is None:
            g["weight"] = gw
        else:
            g["weight"] = np.bincount(mp[mp >= 0], weights=gw.reshape(-1)[mp >= 0], minlength=np.asarray(spec["weights"][l]).size)
        grads[l] = g
        d = d @ dense[l]
    return loss, grads, stats
