"""Losses and schedules of the hot path (reference baler/modules/utils.py:176-323)."""
import math

import numpy as np
import torch

factor = 0.5
min_lr = 1e-6


def mse_sum_loss_l1(model_children, true_data, reconstructed_data, reg_param, validate):
    """sum((recon - true)^2) / n_columns (reference utils.py:195-199).  With validate=False the
    reference adds reg_param * L1 of a second ReLU chain (utils.py:201-209); that branch only exists
    fused inside the training kernel here (`bb_train_hyper.l1`), so it is not available on loose tensors."""
    from .. import engine

    if not validate:
        raise NotImplementedError("the L1 branch runs inside bb_trainer_step (config.l1_in_training)")
    t = true_data.to(device="cuda", dtype=torch.float32).contiguous()
    r = reconstructed_data.to(device="cuda", dtype=torch.float32).contiguous()
    return engine.mse_sum(r, t) / true_data.shape[1], 0, 0


class EarlyStopping:
    """reference utils.py:248-282"""

    def __init__(self, patience, min_delta):
        self.patience, self.min_delta = patience, min_delta
        self.counter, self.best_loss, self.early_stop = 0, None, False

    def __call__(self, train_loss):
        if self.best_loss is None:
            self.best_loss = train_loss
        elif self.best_loss - train_loss > self.min_delta:
            self.best_loss = train_loss
            self.counter = 0
        elif self.best_loss - train_loss < self.min_delta:
            self.counter += 1
            print(f"Early stopping counter {self.counter} of {self.patience}")
            if self.counter >= self.patience:
                print("Early Stopping")
                self.early_stop = True


class LRScheduler:
    """reference utils.py:285-323: ReduceLROnPlateau(mode="min", factor, patience, min_lr), torch's default
    relative threshold 1e-4 and eps 1e-8.  `optimizer` is anything with `param_groups[0]["lr"]` (a torch
    optimizer) or an `lr` attribute (the bb_trainer wrapper of training.py)."""

    def __init__(self, optimizer, patience, min_lr=min_lr, factor=factor):
        self.optimizer, self.patience, self.min_lr, self.factor = optimizer, patience, min_lr, factor
        self.best, self.num_bad_epochs = math.inf, 0

    def _get(self):
        return self.optimizer.param_groups[0]["lr"] if hasattr(self.optimizer, "param_groups") else self.optimizer.lr

    def _set(self, lr):
        if hasattr(self.optimizer, "param_groups"):
            for g in self.optimizer.param_groups:
                g["lr"] = lr
        else:
            self.optimizer.lr = lr

    def __call__(self, train_loss):
        current = float(train_loss)
        if current < self.best * (1.0 - 1e-4):
            self.best, self.num_bad_epochs = current, 0
        else:
            self.num_bad_epochs += 1
        if self.num_bad_epochs > self.patience:
            old = self._get()
            new = max(old * self.factor, self.min_lr)
            if old - new > 1e-8:
                self._set(new)
            self.num_bad_epochs = 0
