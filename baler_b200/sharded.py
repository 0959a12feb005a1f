"""Multi-GPU host logic: one process per GPU (torch.distributed), rows sharded contiguously.

compress / decompress: no collective on the data path; the only exchange is the 2 x n_features
column min / max (all-reduce MIN and MAX) when normalisation features are recomputed from the input
(reference helper.py:500-502 computes them over the whole file).
training: data-parallel; every global batch (reference order, shuffle=False) is cut into `world`
contiguous sub-batches; the flat gradient (+ batch loss in its last slot) is all-reduced with SUM -
not mean - because the reference loss is a sum over rows (utils.py:195-199), then every rank applies
the identical Adam step.
"""
import torch
import torch.distributed as dist


def row_range(n_rows, rank, world):
    """contiguous shard [lo, hi) of n_rows rows for `rank` (first n_rows % world ranks get one extra row)"""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def dp_batch_slices(n_rows, global_batch, rank, world):
    """[(lo, hi)] global row ranges this rank processes, one per global batch (drop_last=False: the last
    batch is ragged and a rank's slice of it may be empty)"""
    out = []
    for b0 in range(0, n_rows, global_batch):
        rows = min(global_batch, n_rows - b0)
        lo, hi = row_range(rows, rank, world)
        out.append((b0 + lo, b0 + hi))
    return out


def combine_minmax_(mn, mx, group=None):
    """in-place global column min / max over all ranks"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return mn, mx


def allreduce_sum_(flat, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


class DataParallelTrainer:
    """engine.Trainer on every rank + SUM all-reduce of the flat gradient between backward and Adam."""

    def __init__(self, trainer, group=None):
        self.trainer, self.group = trainer, group
        self.grads = trainer.grads_view()  # n_params + 1 floats: gradient, then the batch loss
        # AE_Dropout_BN: every rank normalises ITS slice of the batch (per-rank BatchNorm statistics, as torch
        # DistributedDataParallel runs the reference model) and draws its own dropout stream; the running statistics
        # are averaged over ranks at the end of an epoch so that the replicas save the same model.pt
        self.bn_running = trainer.bn_running_views() if getattr(trainer, "_bn", None) is not None else None

    def sync_running_stats(self):
        if self.bn_running is None or not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(self.group)
        if world > 1:
            for t in self.bn_running:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
                t.div_(world)

    def step(self, x_local, hyper):
        """x_local: this rank's contiguous slice of the global batch (may have 0 rows)"""
        if x_local.shape[0] > 0:
            self.trainer.step(x_local, hyper, phase=1)
        else:
            self.grads.zero_()
        allreduce_sum_(self.grads, self.group)
        self.trainer.step(x_local if x_local.shape[0] > 0 else self._dummy(x_local), hyper, phase=2)

    def _dummy(self, x_local):
        return torch.zeros((1, x_local.shape[1]), dtype=x_local.dtype, device=x_local.device)

    def epoch(self, x_local_batches, hyper):
        """x_local_batches: list of this rank's slices, one per global batch; returns the epoch loss"""
        self.trainer.loss_accum.zero_()
        for xb in x_local_batches:
            self.step(xb, hyper)
        self.sync_running_stats()
        return self.trainer.loss_accum.item() / max(len(x_local_batches), 1)
