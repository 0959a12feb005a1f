"""Conv_AE training step (batch 600) on the layer-by-layer trainer: python tools/cfd_train_bench.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baler_b200 import engine, synth  # noqa: E402
from baler_b200.modules import models  # noqa: E402

torch.manual_seed(0)
cm = models.Conv_AE(5, 250)
snaps = synth.cfd_snapshots(600)
snaps = (snaps - snaps.min()) / (snaps.max() - snaps.min())
tb = torch.from_numpy(np.ascontiguousarray(snaps.reshape(-1, 25)[:60000])).cuda()
sp = cm.training_spec(5, 5)
ctr = engine.LayeredTrainer(sp["weights"], sp["biases"], sp["acts"], 600, dims=sp["dims"], w_maps=sp["w_maps"], bn=sp["bn"], loss_columns=1)
steps = int(os.environ.get("STEPS", "100"))
ctr.epoch(tb[:6000], 600, engine.make_hyper(lr=1e-3))
torch.cuda.synchronize()
t0 = time.perf_counter()
loss = ctr.epoch(tb[:600 * steps], 600, engine.make_hyper(lr=1e-3))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("%.1f us / step, loss %.4f" % (1e6 * dt / steps, loss))
