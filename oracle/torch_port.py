"""CPU timing port of the reference path on the reference's own arithmetic engine (torch, CPU, float64).

TEST / BENCH INFRASTRUCTURE - NOT PRODUCT CODE (same rule as baler_oracle.py).  Used by the `cpu_baseline`
and `--impl reference` legs of bench.py: it gives the CPU its best case - vectorised normalisation, the
AE evaluated in cache-sized row blocks into a preallocated output, all torch threads - instead of the
shipped per-batch DataLoader loop with np.concatenate growth (reference helper.py:583-611), which is
several times slower.  Checked against the numpy oracle in tests/test_oracle_golden.py.
Restates reference baler/modules/models.py:141-152 and helper.py:473-616, 619-733.
"""
import numpy as np
import torch
import torch.nn.functional as F

BLOCK = 8192
ENC = ("en1", "en2", "en3", "en4")
DEC = ("de1", "de2", "de3", "de4")


def to_torch(sd):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(torch.float64) for k, v in sd.items()}


def _chain(sd, names, x):
    h = x
    for n in names[:3]:
        h = F.leaky_relu(F.linear(h, sd[n + ".weight"], sd[n + ".bias"]))
    return F.linear(h, sd[names[3] + ".weight"], sd[names[3] + ".bias"])


@torch.no_grad()
def compress(sd, table, block=BLOCK):
    """min/max of this table -> (x - min) / range in float32 -> float64 encode; returns (z float64, [min; range])"""
    x = torch.from_numpy(table)
    mn, mx = x.min(dim=0).values, x.max(dim=0).values
    rg = mx - mn
    out = torch.empty((x.shape[0], sd["en4.bias"].shape[0]), dtype=torch.float64)
    for i in range(0, x.shape[0], block):
        xb = ((x[i:i + block] - mn) / rg).to(torch.float64)
        out[i:i + block] = _chain(sd, ENC, xb)
    return out.numpy(), torch.stack([mn, rg]).numpy()


@torch.no_grad()
def decompress(sd, z, feats, block=BLOCK):
    zt = torch.from_numpy(z)
    mn, rg = torch.from_numpy(feats[0]).to(torch.float64), torch.from_numpy(feats[1]).to(torch.float64)
    out = torch.empty((zt.shape[0], sd["de4.bias"].shape[0]), dtype=torch.float64)
    for i in range(0, zt.shape[0], block):
        out[i:i + block] = _chain(sd, DEC, zt[i:i + block]) * rg + mn
    return out.numpy()


def fit_steps(sd, x_norm, batch, n_steps, lr=1e-3, as_shipped=False):
    """Timing port of the reference's training loop body (training.py:64-97) on torch CPU float64:
    zero_grad -> forward -> sum-MSE / n_cols -> backward -> Adam.step -> loss.item() per batch.
    as_shipped: the batches come out of torch.utils.data.DataLoader(tensor, batch_size, shuffle=False, drop_last=False) as in
    training.py:253-263 (row-by-row indexing + collation per batch) instead of tensor slices.
    Returns (mean loss, seconds)."""
    import time

    if as_shipped:
        from torch.utils.data import DataLoader
        params = {k: v.clone().requires_grad_(True) for k, v in to_torch(sd).items()}
        opt = torch.optim.Adam(list(params.values()), lr=lr)
        x = torch.from_numpy(np.ascontiguousarray(x_norm)).to(torch.float64)[:batch * n_steps]
        total, n = 0.0, 0
        t0 = time.perf_counter()
        for xb in DataLoader(x, batch_size=batch, shuffle=False, drop_last=False):
            opt.zero_grad()
            loss = ((_chain(params, DEC, _chain(params, ENC, xb)) - xb) ** 2).sum() / xb.shape[1]
            loss.backward()
            opt.step()
            total += loss.item()
            n += 1
        return total / max(n, 1), time.perf_counter() - t0

    params = {k: v.clone().requires_grad_(True) for k, v in to_torch(sd).items()}
    opt = torch.optim.Adam(list(params.values()), lr=lr)
    x = torch.from_numpy(np.ascontiguousarray(x_norm)).to(torch.float64)
    total = 0.0
    t0 = time.perf_counter()
    for s in range(n_steps):
        xb = x[(s * batch) % max(len(x) - batch, 1):][:batch]
        opt.zero_grad()
        loss = ((_chain(params, DEC, _chain(params, ENC, xb)) - xb) ** 2).sum() / xb.shape[1]
        loss.backward()
        opt.step()
        total += loss.item()
    return total / n_steps, time.perf_counter() - t0
