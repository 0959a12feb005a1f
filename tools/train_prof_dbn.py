"""where a tensor-core AE_Dropout_BN (or AE) step spends its time: SM-clock stamps of CTA 0: python tools/train_prof_dbn.py [ae]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baler_b200 import _lib, engine, synth  # noqa: E402
from baler_b200.modules import models  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "dbn"
n, bs = 100_000, 512
x = synth.cms_table_device(n, seed=1, device="cuda")
mn, mx = engine.colminmax(x)
xt = engine.normalize_table(x, mn, mx - mn)
torch.manual_seed(0)
if kind == "ae":
    w, b = models.AE(24, 15).linear_tensors()
    tr = engine.Trainer(w, b, 24, 15, bs)
else:
    dm = models.AE_Dropout_BN(24, 15)
    w, b = dm.linear_tensors()
    tr = engine.Trainer(w, b, 24, 15, bs, bn=dm.bn_tensors())
    tr.set_dropout(seed=3)
h = engine.make_hyper(lr=1e-3)
tr.epoch(xt[:51200], bs, h)
_lib.lib().bb_trainer_profile(tr.handle, 50, None)
tr.epoch(xt, bs, h)
st = np.zeros(1024, dtype=np.int64)
_lib.lib().bb_trainer_profile(tr.handle, -1, st.ctypes.data)
t0 = st[0]
print("%s cycles: phase1 %d | barrier %d | phase2 %d | barrier %d | total %d" %
      (kind, st[1] - st[0], st[2] - st[1], st[3] - st[2], st[4] - st[3], st[4] - st[0]))
print("per layer pass and warp: start | tile-major write / mma loop incl. weight wait / epilogue")
for q in range(15):
    rec = st[16 + q * 32: 16 + q * 32 + 32].reshape(8, 4)
    print("  pass %2d:" % q, " ".join("%d|%d/%d/%d" % (r[0] - t0, r[3] - r[0], r[1] - r[3], r[2] - r[1]) for r in rec))
if kind != "ae":
    print("BatchNorm reduction points (cycles from the start of the point's epilogue): packets stored | hop 1 (owned columns "
          "reduced, final packets published) | hop 2 (final packets gathered) | barrier passed")
    for pt in range(8):
        q = st[520 + 8 * pt: 520 + 8 * pt + 6]
        print("  point %d: start %d |" % (pt, q[0] - t0), " ".join(str(int(q[k] - q[0])) for k in (1, 2, 4, 5)))
