"""Where does bb_compress_host / bb_decompress_host spend its time?  (run on the GPU box)"""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from baler_b200 import synth
from baler_b200.modules import models
g = np.load('tests/golden/ae_cms.npz')
m = models.AE(24, 15); m.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('sd/')})
codec = m.eval().codec()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
x = synth.cms_table_device(n)
xh = torch.empty((n, 24), dtype=torch.float32, pin_memory=True); xh.copy_(x)
zh = torch.empty((n, 15), dtype=torch.float32, pin_memory=True)
yh = torch.empty((n, 24), dtype=torch.float32, pin_memory=True)
torch.cuda.synchronize()
def t(f, reps=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
d = torch.empty_like(x)
print('H2D 9.6GB  %.1f GB/s' % (n * 96 / t(lambda: d.copy_(xh, non_blocking=True)) / 1e9))
print('D2H 9.6GB  %.1f GB/s' % (n * 96 / t(lambda: yh.copy_(d, non_blocking=True)) / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(xh, non_blocking=True)
    with torch.cuda.stream(s2): zh.copy_(torch.empty((n, 15), device='cuda') if False else z_dev, non_blocking=True)
z_dev = torch.empty((n, 15), device='cuda')
print('H2D 9.6 + D2H 6.0 concurrently: %.3f s' % t(both))
feats = None
def comp():
    global feats
    _, feats = codec.compress_host(xh.numpy(), recompute_minmax=True, z_dtype=np.float32, out=zh.numpy())
print('compress_host (recompute min/max) %.3f s' % t(comp))
print('compress_host (features given)    %.3f s' % t(lambda: codec.compress_host(xh.numpy(), features=feats, z_dtype=np.float32, out=zh.numpy())))
print('decompress_host                   %.3f s' % t(lambda: codec.decompress_host(zh.numpy(), features=feats, y_dtype=np.float32, out=yh.numpy())))
t0 = time.perf_counter(); b = torch.empty(n * 24, dtype=torch.float32, device='cuda'); torch.cuda.synchronize(); t1 = time.perf_counter(); del b; torch.cuda.synchronize(); print('cudaMalloc+free 9.6 GB via torch: %.4f s' % (t1 - t0))
