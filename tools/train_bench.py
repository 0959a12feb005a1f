"""time one training epoch (600k rows, bs 512) on both arithmetic paths of bb_trainer: python tools/train_bench.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baler_b200 import engine, synth  # noqa: E402
from baler_b200.modules import models  # noqa: E402

n, bs = 600_000, int(os.environ.get("BS", "512"))
x = synth.cms_table_device(n, seed=1, device="cuda")
mn, mx = engine.colminmax(x)
xt = engine.normalize_table(x, mn, mx - mn)
for prec in ("split16", "fp32"):
    torch.manual_seed(0)
    w, b = models.AE(24, 15).linear_tensors()
    tr = engine.Trainer(w, b, 24, 15, bs)
    tr.set_precision(prec)
    h = engine.make_hyper(lr=1e-3)
    tr.epoch(xt[:51200], bs, h)
    torch.cuda.synchronize()
    res = []
    for _ in range(3):
        t0 = time.perf_counter()
        loss = tr.epoch(xt, bs, h)
        torch.cuda.synchronize()
        res.append(time.perf_counter() - t0)
    steps = (n + bs - 1) // bs
    print("%s: %.2f us/step, %.2f M samples/s, epoch loss %.6f, flag %s"
          % (prec, 1e6 * min(res) / steps, n / min(res) / 1e6, loss, tr.range_flag() if prec == "split16" else "-"), flush=True)

# where a split16 step spends its time: SM-clock stamps of CTA 0 during step 100 of an epoch
import ctypes as C  # noqa: E402
import numpy as np  # noqa: E402
from baler_b200 import _lib  # noqa: E402
torch.manual_seed(0)
w, b = models.AE(24, 15).linear_tensors()
tr = engine.Trainer(w, b, 24, 15, bs)
h = engine.make_hyper(lr=1e-3)
tr.epoch(xt[:51200], bs, h)
nch = _lib.lib().bb_trainer_profile(tr.handle, 100, None)
tr.epoch(xt, bs, h)
st = np.zeros(1024, dtype=np.int64)
_lib.lib().bb_trainer_profile(tr.handle, -1, st.ctypes.data)
t0 = st[0]
print("cycles: phase1 %d | barrier %d | phase2 %d | barrier %d | total %d" %
      (st[1] - st[0], st[2] - st[1], st[3] - st[2], st[4] - st[3], st[4] - st[0]))
print("per layer pass and warp: [start after the pass barrier (cycles from step start) | feature-major write / mma loop incl. waiting for weights / epilogue]")
for q in range(15):
    rec = st[16 + q * 32: 16 + q * 32 + 32].reshape(8, 4)
    print("  pass %2d:" % q, " ".join("%d|%d/%d/%d" % (r[0] - t0, r[3] - r[0], r[1] - r[3], r[2] - r[1]) for r in rec))
q = st[900:913]
print("phase 2, CTA 0, item 0 (cycles from item start): group waits / mma ends", [int(v - q[0]) for v in q[1:9]],
      "| k loop done %d | adam done %d | packed %d" % (q[10] - q[0], q[11] - q[0], q[12] - q[0]))
