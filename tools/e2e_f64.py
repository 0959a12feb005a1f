"""bb_compress_host + bb_decompress_host with the reference's file dtypes (float64 latent and reconstruction): python tools/e2e_f64.py"""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from baler_b200 import synth
from baler_b200.modules import models
g = np.load('tests/golden/ae_cms.npz')
m = models.AE(24, 15); m.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('sd/')})
codec = m.eval().codec()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
x = synth.cms_table_device(n)
xh = torch.empty((n, 24), dtype=torch.float32, pin_memory=True); xh.copy_(x)
torch.cuda.synchronize()
xn = xh.numpy()
z64 = np.empty((n, 15), dtype=np.float64)
y64 = np.empty((n, 24), dtype=np.float64)
for it in range(3):
    t0 = time.perf_counter()
    _, feats = codec.compress_host(xn, recompute_minmax=True, z_dtype=np.float64, out=z64)
    t1 = time.perf_counter()
    codec.decompress_host(z64, features=feats, y_dtype=np.float64, out=y64)
    t2 = time.perf_counter()
    print("compress %.3f s, decompress %.3f s -> %.1f M rows/s" % (t1 - t0, t2 - t1, n / (t2 - t0) / 1e6))
z32, _ = codec.compress_host(xn, features=feats, z_dtype=np.float32)
print("widening exact:", np.array_equal(z32.astype(np.float64), z64))
