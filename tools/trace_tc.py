"""SM-clock trace of one tile pipeline of the tcgen05 chain kernel (CTA 0, pipeline 0).
usage: python tools/trace_tc.py [groups]   (groups > 0 forces the table-driven kernel with that many pipelines)"""
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, '.')
from baler_b200 import _lib
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
    dbg = torch.zeros(3 * 16 * 64, dtype=torch.int32, device='cuda')
    for it in range(2):
        rc = fn(codec.handle, decode, xin.data_ptr(), n, out.data_ptr(), 0, -2, dbg.data_ptr(), GROUPS, None)
        assert rc == 0
        torch.cuda.synchronize()
    full = dbg.cpu().numpy().astype(np.int64) & 0xffffffff
    t = full[:1024].reshape(16, 64)
    import os
    pipe = int(os.environ.get('BALER_B200_TRACE_PIPE', '0'))
    ks = full[1024 * (1 + pipe):1024 * (2 + pipe)].reshape(16, 64)
    print('decode' if decode else 'encode', 'tile period (cycles):', np.diff(t[2:12, 0]))
    for lt in (5, 6):
        rel = lambda v: int((v - t[lt, 0]) & 0xffffffff) if v else -1
        e = [rel(v) for v in t[lt, :32]]
        mm = [rel(v) for v in t[lt, 32:]]
        print(' tile', lt, 'EPI: start 0 | next-tile a1 conversion (h=1 warps) from', e[1], 'to', e[2])
        for s in range(5):
            print('   step', s, 'full_d seen', e[3 + 4 * s], 'sub0 done', e[4 + 4 * s], 'sub1 done', e[5 + 4 * s],
                  '| MMA waits', [mm[2 + 5 * s + c] for c in range(4)], 'commit', mm[2 + 5 * s + 4])
        kk = [rel(v) for v in ks[lt, :56]]
        print('   issuer k-steps (wait passed, issued):', ' '.join(f'{kk[2*i]}/{kk[2*i+1]}' for i in range(28)))
        print('   a1: in_full passed', e[29], 'converted', e[30], 'st waited', e[31], '| final: read-wait done', e[28])
        print('   final: ld done', e[25], 'stage written', e[26], 'store issued', e[27], 'tile done', e[24], '| MMA tile start', mm[0], 'a1 seen', mm[1])
