"""Merged issue timeline of the two tile pipelines of chain_tc4_kernel (CTA 0): where is the tensor pipe idle?
usage: python tools/trace_tc_pipes.py   (run on the GPU box)"""
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
n = 148 * 2 * 128 * 16
# pipe cycles of the 28 k-steps (3 MMAs each at max(32, N / 2) + 3.5): encode, decode
COST = {0: [3 * 59.5] * 2 + [3 * 59.5] * 7 + [3 * 51.5] * 2 + [3 * 59.5] * 6 + [3 * 35.5] * 7 + [3 * 32.6] * 4,
        1: [3 * 35.5] * 1 + [3 * 59.5] * 4 + [3 * 59.5] * 7 + [3 * 32.6] * 7 + [3 * 51.5] * 7 + [3 * 32.6] * 6}
for decode, dim_in, dim_out in ((0, 24, 15), (1, 15, 24)):
    xin = torch.rand((n, dim_in), device='cuda')
    out = torch.empty((n, dim_out), device='cuda')
    dbg = torch.zeros(3 * 1024, dtype=torch.int32, device='cuda')
    for it in range(2):
        assert fn(codec.handle, decode, xin.data_ptr(), n, out.data_ptr(), 0, -2, dbg.data_ptr(), 0, None) == 0
        torch.cuda.synchronize()
    full = dbg.cpu().numpy().astype(np.int64) & 0xffffffff
    ev = []  # (issued clock, wait-passed clock, pipe, tile, kstep)
    nks = len(COST[decode])
    for pipe in (0, 1):
        ks = full[1024 * (1 + pipe):1024 * (2 + pipe)].reshape(16, 64)
        for lt in range(4, 12):
            for k in range(nks):
                ev.append((int(ks[lt, 2 * k + 1]), int(ks[lt, 2 * k]), pipe, lt, k))
    ev.sort()
    t0, t1 = ev[0][1], ev[-1][0]
    busy = sum(COST[decode][e[4]] for e in ev)
    print('decode' if decode else 'encode', 'window %d cycles, %d k-steps, pipe work %.0f cycles = %.1f %% busy' % (t1 - t0, len(ev), busy, 100 * busy / (t1 - t0)))
    # time an issuer spent waiting for operands (wait passed - previous issue of the same pipe) vs issuing (issued - wait passed)
    for pipe in (0, 1):
        pe = sorted([e for e in ev if e[2] == pipe])
        wait = sum(max(0, pe[i][1] - pe[i - 1][0]) for i in range(1, len(pe)))
        issue = sum(e[0] - e[1] for e in pe)
        print('  pipeline %d: waiting for operands %d cycles, issuing (incl. pipe back-pressure) %d cycles' % (pipe, wait, issue))
    # idle gaps of the pipe: both issuers waiting.  approximate pipe-free time: issue end + cost of that k-step
    free, idle, gaps = ev[0][1], 0, []
    for issued, passed, pipe, lt, k in ev:
        start = issued - 0  # the k-step's MMAs entered the pipe by `issued`
        begin = max(free, passed)
        if passed > free:
            idle += passed - free
            gaps.append((passed - free, pipe, lt, k))
        free = begin + COST[decode][k]
    print('  estimated pipe idle %d cycles (%.1f %%); largest gaps (cycles, pipeline, tile, k-step): %s' % (idle, 100 * idle / (t1 - t0), sorted(gaps, reverse=True)[:12]))
    bykstep = {}
    for gsz, pipe, lt, k in gaps:
        bykstep[k] = bykstep.get(k, 0) + gsz
    print('  idle by k-step index:', sorted(bykstep.items(), key=lambda kv: -kv[1])[:10])
