"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in baler_b200/sharded.py: row sharding, the
min/max exchange of sharded compress, and SUM (not mean) gradient all-reduce of data-parallel training.
The oracle stands in for the per-rank CUDA compute; the collectives and the slicing are the code under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, rel_max
from baler_b200 import sharded, synth


def test_row_range_partitions_exactly():
    for n in (0, 1, 7, 100, 1001, 600_000):
        for world in (1, 2, 3, 8):
            cuts = [sharded.row_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_dp_batch_slices_preserve_reference_order():
    n, gb = 600_000, 1024  # T600k with 512 rows per GPU on 2 GPUs; last global batch is ragged (960 rows)
    per_rank = [sharded.dp_batch_slices(n, gb, r, 2) for r in range(2)]
    assert len(per_rank[0]) == len(per_rank[1]) == (n + gb - 1) // gb
    for b, (s0, s1) in enumerate(zip(*per_rank)):
        assert s0[0] == b * gb and s0[1] == s1[0] and s1[1] == min(n, (b + 1) * gb)
    # a rank's slice of a tiny last batch may be empty
    tail = [sharded.dp_batch_slices(1025, 1024, r, 8)[-1] for r in range(8)]
    assert sum(hi - lo for lo, hi in tail) == 1 and sum(1 for lo, hi in tail if hi == lo) == 7


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import baler_oracle as orc

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # --- sharded compress: local column min/max -> exchange -> identical global features on every rank
        table = synth.cms_table(10_001, seed=5)
        lo, hi = sharded.row_range(len(table), rank, world)
        mn, mx = torch.from_numpy(table[lo:hi].min(0)), torch.from_numpy(table[lo:hi].max(0))
        sharded.combine_minmax_(mn, mx)
        feats = np.stack([mn.numpy(), (mx - mn).numpy()])
        assert np.array_equal(feats, orc.find_minmax(table))
        # --- data-parallel step: SUM all-reduce of [grads | loss] == single-process full batch
        g = np.load(os.path.join(GOLDEN, "ae_train.npz"))
        sd0 = {k[4:]: g[k] for k in g.files if k.startswith("sd0/")}
        x = np.ascontiguousarray(g["x_norm"][:1000]).astype(np.float64)
        (s_lo, s_hi), = sharded.dp_batch_slices(len(x), len(x), rank, world)
        loss, _, _, grads = orc.ae_loss_and_grads(sd0, x[s_lo:s_hi])
        keys = sorted(grads)
        flat = torch.from_numpy(np.concatenate([grads[k].ravel() for k in keys] + [[loss]]))
        sharded.allreduce_sum_(flat)
        full_loss, _, _, full = orc.ae_loss_and_grads(sd0, x)
        ref = np.concatenate([full[k].ravel() for k in keys] + [[full_loss]])
        assert rel_max(flat.numpy(), ref) < 1e-12
        # every rank applies the identical Adam step
        opt = orc.Adam({k: sd0[k].copy() for k in keys})
        off, summed = 0, {}
        for k in keys:
            summed[k] = flat.numpy()[off:off + grads[k].size].reshape(grads[k].shape)
            off += grads[k].size
        opt.step(summed)
        np.save(os.path.join(out_dir, "params_rank%d.npy" % rank), np.concatenate([opt.params[k].ravel() for k in keys]))
        # --- sharded CLI compress / decompress: global [min; range] from row shards, rows gathered on rank 0 in row order
        for n in (10_001, 1):  # 1 row: rank 1's shard is empty
            t = synth.cms_table(max(n, 2), seed=6)[:n]
            lo, hi = sharded.row_range(n, rank, world)
            f = sharded.global_minmax(t[lo:hi])
            assert f.dtype == np.float32 and np.array_equal(f, orc.find_minmax(t))
            full = sharded.gather_rows_to_rank0(np.ascontiguousarray(t[lo:hi, :15]).astype(np.float64), n, chunk_rows=1000)
            assert (full is None) == (rank != 0)
            if rank == 0:
                assert full.dtype == np.float64 and np.array_equal(full, t[:, :15].astype(np.float64))
        # --- a float64 file whose offset dwarfs its spread (data_processing.F32_OFFSET_LIMIT): every rank reaches the same
        # decision from the row sample and gets the exact float64 statistics of the WHOLE table from the shard exchange,
        # the way helper.compress wires it under torchrun
        from baler_b200.modules import data_processing as dp
        t64 = synth.cms_table(10_001, seed=8).astype(np.float64)
        t64[:, 3] = 2.0e9 + np.arange(len(t64)) % 701
        lo, hi = sharded.row_range(len(t64), rank, world)

        def shard_stats(flat):
            f = sharded.global_minmax(flat[lo:hi])
            return f[0], f[0] + f[1]

        st = dp.float64_stats(t64, shard_stats)
        ref64 = orc.find_minmax(t64)
        assert st is not None and np.array_equal(st[0], ref64[0]) and np.allclose(st[1], ref64[1], rtol=0, atol=1e-6)
        assert np.allclose(dp.normalize_float64_host(t64, *st), orc.normalize(t64), rtol=0, atol=1e-9)
        assert dp.float64_stats(synth.cms_table(10_001, seed=8).astype(np.float64), shard_stats) is None  # no exchange needed
        # --- sharded.DataParallelTrainer itself (the class the GPU trainers run under) on a CPU stand-in trainer built on the
        # oracle: phase 1 fills [grads | loss], SUM all-reduce, phase 2 applies Adam; a rank with an empty slice of the last
        # batch contributes zeros; BatchNorm running statistics (one tensor, as the layer-by-layer trainer exposes them)
        # are averaged over ranks at the end of the epoch
        class OracleTrainer:
            def __init__(self):
                self.opt = orc.Adam({k: sd0[k].copy() for k in keys})
                self.n = sum(sd0[k].size for k in keys)
                self.grads = torch.zeros(self.n + 1, dtype=torch.float64)
                self.loss_accum = torch.zeros(1, dtype=torch.float64)
                self._bn = True
                self.running = torch.full((6,), float(rank + 1), dtype=torch.float64)

            def grads_view(self):
                return self.grads

            def bn_running_views(self):
                return (self.running,)

            def step(self, xb, hyper, phase):
                if phase == 1:
                    l, _, _, gr = orc.ae_loss_and_grads(self.opt.params, xb.numpy())
                    self.grads.copy_(torch.from_numpy(np.concatenate([gr[k].ravel() for k in keys] + [[l]])))
                else:
                    off, summed = 0, {}
                    for k in keys:
                        summed[k] = self.grads.numpy()[off:off + sd0[k].size].reshape(sd0[k].shape)
                        off += sd0[k].size
                    self.opt.step(summed)
                    self.loss_accum += self.grads[-1]

        tr = OracleTrainer()
        dp = sharded.DataParallelTrainer(tr)
        xt = torch.from_numpy(x[:513])  # global batch 512: the last batch is ONE row, rank 1's slice of it is empty
        slices = sharded.dp_batch_slices(513, 512, rank, world)
        assert (slices[-1][1] - slices[-1][0] == 0) == (rank == 1)
        epoch_loss = dp.epoch([xt[a:b] for a, b in slices], None)
        single = orc.Adam({k: sd0[k].copy() for k in keys})
        ref_loss = 0.0
        for a, b in ((0, 512), (512, 513)):
            l, _, _, gr = orc.ae_loss_and_grads(single.params, x[a:b])
            single.step(gr)
            ref_loss += l
        assert abs(epoch_loss - ref_loss / 2) < 1e-9 * ref_loss
        for k in keys:
            assert rel_max(tr.opt.params[k], single.params[k]) < 1e-9, k
        assert torch.equal(tr.running, torch.full((6,), 1.5, dtype=torch.float64))
        # --- variable-length host lists (error-bounded-delta hits with global row numbers) collected in rank order
        hits = (np.arange(rank * 5, rank * 5 + 3 + rank, dtype=np.int64), np.full(3 + rank, rank, dtype=np.int64),
                np.full(3 + rank, 0.5 * rank, dtype=np.float16))
        got = sharded.gather_objects_to_rank0(hits)
        assert (got is None) == (rank != 0)
        if rank == 0:
            assert [len(h[0]) for h in got] == [3, 4] and got[1][2].dtype == np.float16
            assert np.array_equal(np.concatenate([h[0] for h in got]), [0, 1, 2, 5, 6, 7, 8])
        # --- every replica starts from rank 0's initial model (each process draws its own random weights otherwise)
        from baler_b200.modules import models, training
        torch.manual_seed(100 + rank)
        m = models.AE_Dropout_BN(24, 15)
        before = torch.cat([v.double().ravel() for v in m.state_dict().values()])
        training.broadcast_initial_state(m)
        after = torch.cat([v.double().ravel() for v in m.state_dict().values()])
        assert (rank == 0) == bool(torch.equal(before, after))
        np.save(os.path.join(out_dir, "init_rank%d.npy" % rank), after.numpy())
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    p0, p1 = np.load(tmp_path / "params_rank0.npy"), np.load(tmp_path / "params_rank1.npy")
    assert np.array_equal(p0, p1)  # replicated optimizer state stays bit-identical across ranks
    assert np.array_equal(np.load(tmp_path / "init_rank0.npy"), np.load(tmp_path / "init_rank1.npy"))
