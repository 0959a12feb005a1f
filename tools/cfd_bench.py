"""Conv_AE 5x5 -> 250 encode / decode on the layered GEMM path, both precisions: python tools/cfd_bench.py [blocks]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baler_b200 import synth  # noqa: E402
from baler_b200.modules import models  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
torch.manual_seed(0)
cm = models.Conv_AE(5, 250).eval()
snaps = synth.cfd_snapshots((nb + 99) // 100)
snaps = (snaps - snaps.min()) / (snaps.max() - snaps.min())
blocks = torch.from_numpy(np.ascontiguousarray(snaps.reshape(-1, 1, 5, 5)[:nb])).cuda()
for prec in ("auto", "fp32"):
    z = cm.encode(blocks, precision=prec); y = cm.decode(z, precision=prec)
    del z, y
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record(); z = cm.encode(blocks, precision=prec); ev[1].record(); y = cm.decode(z, precision=prec); ev[2].record()
    torch.cuda.synchronize()
    te, td = ev[0].elapsed_time(ev[1]) * 1e-3, ev[1].elapsed_time(ev[2]) * 1e-3
    print("%s: encode %.1f M blocks/s (%.1f TFLOP/s), decode %.1f M blocks/s (%.1f TFLOP/s)"
          % (prec, nb / te / 1e6, nb * 1593216 / te / 1e12, nb / td / 1e6, nb * 1593216 / td / 1e12))
