"""Training loop of the hot path (reference baler/modules/training.py): same `fit / validate / train`
entry points; the batch loop runs on the GPU (bb_trainer_epoch) without a host sync per batch."""
import os
import time

import numpy as np
import torch

from .. import engine, sharded
from . import utils


dist_env = sharded.dist_env
FUSED_TRAINER_MAX_FEATURES = 100  # bb_trainer stages whole weight matrices in shared memory (widest: 200 x 100)


def broadcast_initial_state(model):
    """data-parallel runs start every replica from rank 0's freshly initialised model (what DistributedDataParallel does
    at construction): each process draws its own random initial weights otherwise"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"  # gloo: the CPU tests of the host logic
    sd = model.state_dict()
    for k in sd:
        t = sd[k].to(dev)
        dist.broadcast(t, src=0)
        sd[k] = t.cpu()
    model.load_state_dict(sd)


class DeviceBatches:
    """Stands in for the reference's DataLoader(shuffle=False, drop_last=False) (training.py:253-263):
    the whole (normalised, float32) table resident in HBM plus the batch size."""

    def __init__(self, data, batch_size):
        self.data, self.batch_size = data, int(batch_size)

    def __len__(self):
        return (self.data.shape[0] + self.batch_size - 1) // self.batch_size


class DeviceAdam:
    """Stands in for torch.optim.Adam(model.parameters(), lr) (training.py:266): Adam state lives in
    the bb_trainer; `lr` is what LRScheduler adjusts."""

    def __init__(self, model, lr, max_batch, l1=False, reg_param=0.0, dropout_seed=0, block_hw=None, swae=False):
        self.rank, self.world = dist_env()
        broadcast_initial_state(model)
        self.model = model
        self.has_bn = hasattr(model, "bn_tensors")
        self.conv = hasattr(model, "training_spec")
        self.trainer, self.steps = None, 0
        self.lr, self.l1, self.reg_param, self.swae = lr, l1, reg_param, swae
        if swae and (self.conv or self.has_bn or self.world > 1):
            # upstream encodes twice per step with this loss (training.py:66-72): BatchNorm / dropout state would move twice
            raise NotImplementedError("loss_function_swae: dense models without BatchNorm / dropout, one process")
        if self.conv:
            # Conv_AE: (transposed) convolutions as weight-sharing dense layers, BatchNorm2d on batch statistics; the loss
            # divisor is true_data.shape[1] = 1 channel (utils.py:197)
            sp = model.training_spec(*block_hw)
            self.trainer = engine.LayeredTrainer(sp["weights"], sp["biases"], sp["acts"], max_batch, dims=sp["dims"],
                                                 w_maps=sp["w_maps"], bn=sp["bn"], loss_columns=1)
            self.dp = sharded.DataParallelTrainer(self.trainer) if self.world > 1 else None
            return
        w, b = model.linear_tensors()
        if not swae and (self.has_bn or model.n_features <= FUSED_TRAINER_MAX_FEATURES):
            try:
                self.trainer = engine.Trainer(w, b, model.n_features, model.z_dim, max_batch,
                                              bn=model.bn_tensors() if self.has_bn else None)
            except Exception:
                if self.has_bn:
                    raise
        if self.trainer is None:
            # wide rows (CFD_dense_AE on flattened 2-D snapshots): the layer-by-layer GEMM trainer
            self.trainer = engine.LayeredTrainer(w, b, ["leaky", "leaky", "leaky", "none"] * 2, max_batch)
        self.dp = sharded.DataParallelTrainer(self.trainer, fused=not l1) if self.world > 1 else None
        if self.has_bn:
            # fused data parallel: one dropout stream keyed by the global batch row, the same on every rank; otherwise every
            # rank draws its own
            fused = self.dp is not None and self.dp.fused
            self.trainer.set_dropout(seed=dropout_seed + (0 if fused else self.rank))

    def hyper(self, world_size=None):
        return engine.make_hyper(lr=self.lr, reg_param=self.reg_param, l1=self.l1,
                                 world_size=self.world if world_size is None else world_size)

    def epoch(self, data, batch_size):
        """one pass over `data` in the reference's batch order; data-parallel when launched with several ranks"""
        self.steps += (data.shape[0] + batch_size - 1) // batch_size
        if self.dp is None:
            return self.trainer.epoch(data, batch_size, self.hyper())
        return self.dp.epoch_table(data, batch_size, self.hyper(), self.rank, self.world)

    SWAE_PROJECTIONS, SWAE_REG_WEIGHT = 2000, 100.0  # defaults of utils.loss_function_swae (utils.py:27-36)

    def epoch_swae(self, data, batch_size, latent_dim):
        """one pass with config.custom_loss_function = "loss_function_swae" (training.py:70-78).  The two random inputs of
        utils.compute_swd are drawn here from torch's global CPU generator in the reference's order - randn_like(z), then
        randn(num_projections, latent_dim) normalised per row (utils.py:57-90) - in float32 (upstream cannot run this loss on
        the float64 AE: its float32 projections meet a double latent), and handed to the device step."""
        tr = self.trainer
        tr.loss_accum.zero_()
        n_batches = 0
        for r0 in range(0, data.shape[0], batch_size):
            xb = data[r0:r0 + batch_size]
            prior = torch.randn((xb.shape[0], latent_dim), dtype=torch.float32)
            proj = torch.randn(self.SWAE_PROJECTIONS, latent_dim)
            proj = proj / proj.norm(dim=1).view(-1, 1)
            tr.step_swae(xb, self.hyper(), prior.cuda(), proj.cuda(), latent_layer=3, reg_weight=self.SWAE_REG_WEIGHT)
            n_batches += 1
        self.steps += n_batches
        return tr.loss_accum.item() / n_batches

    def sync_model(self):
        if self.conv:
            w, b = self.trainer.get_params()
            self.model.load_trained(w, b, self.trainer.get_bn(), self.steps)
            self.steps = 0
            return
        w, b = self.trainer.get_params()
        self.model.set_linear_tensors(w, b)
        if self.has_bn:
            self.model.set_bn_tensors(self.trainer.get_bn())


def fit(config, model, train_dl, model_children, regular_param, optimizer, latent_dim, RHO, l1, n_dimensions):
    """One epoch (reference training.py:31-101).  Returns (epoch_loss, mse_loss, l1_loss, model); the
    reference always evaluates the loss with validate=True, i.e. without the L1 term -
    the L1 branch is opt-in through `config.l1_in_training`."""
    print("### Beginning Training")
    model.train()
    if hasattr(config, "custom_loss_function") and config.custom_loss_function == "loss_function_swae":
        epoch_loss = optimizer.epoch_swae(train_dl.data, train_dl.batch_size, latent_dim)
    else:
        epoch_loss = optimizer.epoch(train_dl.data, train_dl.batch_size)
    print(f"# Finished. Training Loss: {epoch_loss:.6f}")
    return epoch_loss, epoch_loss, 0, model


def validate(model, test_dl, model_children, reg_param, optimizer=None):
    """reference training.py:104-137: eval-mode forward + sum-MSE / n_columns, mean over batches"""
    print("### Beginning Validating")
    model.eval()
    epoch_loss = optimizer.trainer.validate(test_dl.data, test_dl.batch_size)
    print(f"# Finished. Validation Loss: {epoch_loss:.6f}")
    return epoch_loss


def _to_device_table(data, config):
    arr = np.asarray(data)
    if config.data_dimension == 2:
        if config.model_type == "convolutional" and config.model_name != "Conv_AE":
            raise NotImplementedError("convolutional models other than Conv_AE: see DESIGN.md (out of scope)")
        # dense: (N, H * W) rows (training.py:195-204); Conv_AE: (N, 1, H, W) batches (:222-228), the same memory
        arr = arr.reshape(arr.shape[0], arr.shape[1] * arr.shape[2])
    return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).cuda()


def train(model, variables, train_data, test_data, project_path, config):
    """reference training.py:150-348.  Same side effects: `loss_data.npy` ([train; val] per epoch),
    optional `model_{epoch}.pt`, `activations.npy`; returns the trained model."""
    from . import helper

    dist_env()  # under torchrun: pick this rank's GPU before anything is allocated on a device
    if config.deterministic_algorithm:
        import random

        random.seed(0)
        torch.manual_seed(0)
        np.random.seed(0)
    bs = config.batch_size
    train_ds = _to_device_table(train_data, config)
    valid_ds = train_ds if test_data is train_data else _to_device_table(test_data, config)
    train_dl, valid_dl = DeviceBatches(train_ds, bs), DeviceBatches(valid_ds, bs)
    model_children = list(model.children())
    conv = config.data_dimension == 2 and config.model_type == "convolutional"
    optimizer = DeviceAdam(model, config.lr, max_batch=bs, l1=bool(getattr(config, "l1_in_training", False)),
                           reg_param=config.reg_param, dropout_seed=int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF,
                           block_hw=tuple(np.asarray(train_data).shape[1:3]) if conv else None,
                           swae=getattr(config, "custom_loss_function", None) == "loss_function_swae")
    early_stopping = utils.EarlyStopping(config.early_stopping_patience, config.min_delta) if config.early_stopping else None
    lr_scheduler = utils.LRScheduler(optimizer, config.lr_scheduler_patience) if config.lr_scheduler else None
    train_loss, val_loss = [], []
    start = time.time()
    for epoch in range(config.epochs):
        print(f"Epoch {epoch + 1} of {config.epochs}")
        train_epoch_loss, _, _, _ = fit(config, model, train_dl, model_children, config.reg_param, optimizer,
                                        getattr(config, "latent_space_size", None), config.RHO, config.l1,
                                        config.data_dimension)
        train_loss.append(train_epoch_loss)
        if config.test_size:
            val_epoch_loss = validate(model, valid_dl, model_children, config.reg_param, optimizer)
        else:
            val_epoch_loss = train_epoch_loss
        val_loss.append(val_epoch_loss)
        if lr_scheduler:
            lr_scheduler(val_epoch_loss)
        if early_stopping:
            early_stopping(val_epoch_loss)
            if early_stopping.early_stop:
                break
        if config.intermittent_model_saving and epoch % config.intermittent_saving_patience == 0:
            optimizer.sync_model()
            if optimizer.rank == 0:
                helper.model_saver(model, os.path.join(project_path, f"model_{epoch}.pt"))
    end = time.time()
    optimizer.sync_model()
    if optimizer.rank == 0:  # replicas are identical: one writer
        if getattr(config, "activation_extraction", False):
            np.save(os.path.join(project_path, "activations.npy"), optimizer.trainer.activation_means())
        np.save(os.path.join(project_path, "loss_data.npy"), np.array([train_loss, val_loss]))
        if conv:
            # reference training.py:344-346: the conv-stack output shape of the LAST forward, which is validate()'s last
            # batch when test_size > 0, else the last training batch; decompress reads its batch size back (helper.py:650-688)
            last = valid_dl if config.test_size else train_dl
            rows = last.data.shape[0] - (last.data.shape[0] - 1) // bs * bs
            model.set_final_layer_dims(torch.Size((rows,) + tuple(model._conv_out)))
            np.save(os.path.join(project_path, "final_layer.npy"), np.array(model.get_final_layer_dims()))
    print(f"{(end - start) / 60:.3} minutes")
    return model
