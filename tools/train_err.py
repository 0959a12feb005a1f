"""error of one training step against the float64 oracle, per layer and per gradient tensor, both arithmetic paths"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, rel_l2, rel_max, sub_sd  # noqa: E402
from test_gpu_train_tc import make_trainer, reference_pass  # noqa: E402
from baler_b200 import engine  # noqa: E402
from baler_b200.modules import models  # noqa: E402
from oracle import baler_oracle as orc  # noqa: E402

NAMES = models.AE.names
g = load_golden("ae_train.npz")
sd0 = sub_sd(g, "sd0")
x = g["x_norm"][:512]
acts, dz = reference_pass(sd0, x.astype(np.float64))
loss, _, _, grads = orc.ae_loss_and_grads(sd0, x.astype(np.float64))
for prec in ("split16", "fp32"):
    tr = make_trainer(sd0, precision=prec)
    tr.step(torch.from_numpy(x).cuda(), engine.make_hyper(), phase=1)
    got = tr.grads_view().cpu().numpy()
    print(prec, "loss rel err %.2e" % (abs(got[-1] - loss) / loss))
    off = 0
    for n in NAMES:
        for part in (".weight", ".bias"):
            r = grads[n + part].ravel()
            gg = got[off:off + r.size]
            off += r.size
            print("   %-12s max %.2e  l2 %.2e" % (n + part, rel_max(gg, r), rel_l2(gg, r)))
    if prec == "split16":
        for l in range(8):
            xl = tr.debug_layer(0, l, 512)
            zl = tr.debug_layer(1, l, 512)
            print("   X%d max %.2e l2 %.2e | dZ%d max %.2e l2 %.2e" % (l, rel_max(xl[:-1].T, acts[l]), rel_l2(xl[:-1].T, acts[l]),
                                                                    l, rel_max(zl.T, dz[l]), rel_l2(zl.T, dz[l])))
