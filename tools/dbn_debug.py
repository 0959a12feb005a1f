"""layer-by-layer comparison of the tensor-core AE_Dropout_BN step with the oracle (injected masks): python tools/dbn_debug.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baler_b200 import engine  # noqa: E402
from baler_b200.modules import models  # noqa: E402
from oracle import baler_oracle as orc  # noqa: E402

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ae_dbn.npz"))
sd0 = {k[4:]: np.asarray(g[k], order="C") for k in g.files if k.startswith("sd0/")}
LIN = models.AE_Dropout_BN.enc_names + models.AE_Dropout_BN.dec_names
BN = models.AE_Dropout_BN.bn_names
bn = {k: [sd0[b + "." + k] for b in BN] for k in ("weight", "bias", "running_mean", "running_var")}
bn["num_batches_tracked"] = [int(sd0[b + ".num_batches_tracked"]) for b in BN]
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 512
tr = engine.Trainer([sd0[n + ".weight"] for n in LIN], [sd0[n + ".bias"] for n in LIN], 24, 15, 512, bn=bn)
masks = [g["mask%d" % i].astype(np.uint8)[:rows] for i in range(4)]
tr.set_dropout(masks=[torch.from_numpy(np.ascontiguousarray(m)).cuda() for m in masks])
x = g["x_norm"][:rows]
from baler_b200 import _lib
_lib.lib().bb_trainer_profile(tr.handle, 0, None)
tr.step(torch.from_numpy(x).cuda(), engine.make_hyper(lr=1e-3), phase=1)
st = np.zeros(1024, dtype=np.int64)
_lib.lib().bb_trainer_profile(tr.handle, -1, st.ctypes.data)
sd = {k: np.asarray(v, dtype=np.float64) for k, v in sd0.items()}
# oracle internals
h = x.astype(np.float64)
acts = [h]
for i, name in enumerate(orc.DBN_ENC):
    a = orc.linear(h, sd[name + ".weight"], sd[name + ".bias"]) * masks[i] / (1.0 - orc.DROPOUT_P[i])
    h = orc.leaky_relu(a)
    acts.append(h)
for i in range(3):
    a = orc.leaky_relu(orc.linear(h, sd[orc.DBN_DEC[i] + ".weight"], sd[orc.DBN_DEC[i] + ".bias"]))
    xh = (a - a.mean(0)) / np.sqrt(a.var(0) + 1e-5)
    h = xh * sd[BN[i] + ".weight"] + sd[BN[i] + ".bias"]
    acts.append(h)
for l in range(8):
    got = tr.debug_layer(0, l, rows)[:-1].T
    ref = acts[l]
    print("X_%d: max |diff| %.3e (max |ref| %.3e)" % (l, np.abs(got - ref).max(), np.abs(ref).max()))
loss, _, grads, _ = orc.dbn_train_step(sd, x.astype(np.float64), [m.astype(np.float64) for m in masks])
print("loss", tr.grads_view()[-1].item(), loss)
got = tr.debug_layer(0, 6, rows)[:-1].T
err = np.abs(got - acts[6]).max(axis=0)
print("X_6 per-column max error, columns 0..99:")
print(np.array2string(err, precision=1, max_line_width=200))

q = st[800:900].view(np.uint32).reshape(100, 2).copy()
mean_k, m2_k = q[:, 0].view(np.float32), q[:, 1].view(np.float32)
hh = acts[5]
a_ref = orc.leaky_relu(orc.linear(hh, sd[orc.DBN_DEC[1] + ".weight"], sd[orc.DBN_DEC[1] + ".bias"]))
print("mean kernel vs ref (cols 88..99):", mean_k[88:], a_ref.mean(0)[88:])
print("M2 kernel vs ref (cols 88..99):", m2_k[88:], (a_ref.var(0) * rows)[88:])
