import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, '.')
from baler_b200 import _lib, synth
from baler_b200.modules import models
g = np.load('tests/golden/ae_cms.npz')
m = models.AE(24, 15); m.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('sd/')})
codec = m.eval().codec()
fn = _lib.lib().bb_debug_tc_chain
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
GROUPS = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = 148 * 2 * 128 * 16
x = torch.rand((n, 24), device='cuda')
for decode, dim_in, dim_out in ((0, 24, 15), (1, 15, 24)):
    xin = x if not decode else torch.rand((n, 15), device='cuda')
    out = torch.empty((n, dim_out), device='cuda')
    dbg = torch.zeros(16 * 64, dtype=torch.int32, device='cuda')
    for it in range(2):
        rc = fn(codec.handle, decode, xin.data_ptr(), n, out.data_ptr(), 0, -2, dbg.data_ptr(), GROUPS, None)
        assert rc == 0
        torch.cuda.synchronize()
    t = dbg.cpu().numpy().astype(np.int64).reshape(16, 64) & 0xffffffff
    base = t[4, 0]
    print('decode' if decode else 'encode', 'tile period (cycles):', np.diff(t[2:12, 0]))
    for lt in (5, 6):
        e = (t[lt, :32] - t[lt, 0]) & 0xffffffff
        mm = (t[lt, 32:] - t[lt, 0]) & 0xffffffff
        print(' tile', lt, 'EPI: start 0 | in ready', e[1], '| a1 arrive', e[2])
        for s in range(5):
            print('   step', s, 'full_d seen', e[3 + 4 * s], 'chunkA done', e[4 + 4 * s], 'chunkB done', e[5 + 4 * s],
                  '| MMA waits', [int(mm[2 + 5 * s + c]) for c in range(4)], 'commit', mm[2 + 5 * s + 4])
        print('   final done', e[24], 'after bar', e[25], '| MMA tile start', mm[0], 'a1 seen', mm[1])
