"""a short training epoch for ncu: python tools/train_prof.py [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baler_b200 import engine, synth  # noqa: E402
from baler_b200.modules import models  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
x = synth.cms_table_device(512 * steps, seed=1, device="cuda")
mn, mx = engine.colminmax(x)
xt = engine.normalize_table(x, mn, mx - mn)
torch.manual_seed(0)
w, b = models.AE(24, 15).linear_tensors()
tr = engine.Trainer(w, b, 24, 15, 512)
tr.set_precision(os.environ.get("PREC", "split16"))
h = engine.make_hyper(lr=1e-3)
for _ in range(2):
    print(tr.epoch(xt, 512, h))
torch.cuda.synchronize()
