"""Model classes of the hot path with the reference's model protocol (reference
baler/modules/models.py): ctor `(n_features, z_dim)`, `encode / decode / forward`, `train / eval`,
`to`, `parameters`, `children`, `state_dict / load_state_dict(strict=False)`.

They are NOT torch.nn.Modules and do no torch math: the state dict (same keys, shapes and dtypes as
the reference, so `model.pt` files are interchangeable) lives on the host; `encode` / `decode`
run the fused CUDA chain of libbaler_b200 and training runs in `training.train` on a `bb_trainer`.
Initial weights are drawn with torch's own nn.Linear initialiser in the reference's construction
order, so `torch.manual_seed(s); AE(24, 15)` starts from the reference's weights.
"""
from collections import OrderedDict

import numpy as np
import torch

from .. import engine

BN_EPS = 1e-5


def _as_cuda_f32(x):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    engine.require_cuda()
    return x.to(device="cuda", dtype=torch.float32).contiguous()


class _Layer:
    """What `model.children()` yields: a handle on one Linear's tensors (reference: nn.Linear)."""

    def __init__(self, sd, prefix):
        self._sd, self._prefix = sd, prefix

    @property
    def weight(self):
        return self._sd[self._prefix + ".weight"]

    @property
    def bias(self):
        return self._sd[self._prefix + ".bias"]


class _DenseBase:
    dtype = torch.float64
    hidden = (200, 100, 50)

    def __init__(self, n_features, z_dim, *args, **kwargs):
        self.n_features, self.z_dim = n_features, z_dim
        self.training = True
        self.activations = {}
        self._sd = OrderedDict()
        self._codec = None
        self._device = torch.device("cuda")
        self._build()

    # -- construction helpers
    def _linear(self, name, n_in, n_out):
        lin = torch.nn.Linear(n_in, n_out, dtype=self.dtype)  # consumes the RNG exactly like the reference
        self._sd[name + ".weight"] = lin.weight.detach().clone()
        self._sd[name + ".bias"] = lin.bias.detach().clone()

    # -- nn.Module-like protocol
    def state_dict(self):
        return OrderedDict((k, v.clone()) for k, v in self._sd.items())

    def load_state_dict(self, sd, strict=False):
        missing = [k for k in self._sd if k not in sd]
        unexpected = [k for k in sd if k not in self._sd]
        if strict and (missing or unexpected):
            raise RuntimeError("missing keys %s, unexpected keys %s" % (missing, unexpected))
        for k, v in sd.items():
            if k in self._sd:
                v = torch.as_tensor(v)
                if tuple(v.shape) != tuple(self._sd[k].shape):
                    raise RuntimeError("size mismatch for %s: %s vs %s" % (k, tuple(v.shape), tuple(self._sd[k].shape)))
                self._sd[k] = v.detach().to("cpu", self._sd[k].dtype).clone()
        self._codec = None
        return missing, unexpected

    def parameters(self):
        return [v for k, v in self._sd.items() if k.endswith(("weight", "bias"))]

    def named_parameters(self):
        return [(k, v) for k, v in self._sd.items() if k.endswith(("weight", "bias"))]

    def train(self, mode=True):
        self.training = bool(mode)
        return self

    def eval(self):
        return self.train(False)

    def to(self, device=None, *a, **k):
        return self

    def float(self):
        return self

    @property
    def type(self):
        return type(self).__name__

    def __call__(self, x):
        return self.forward(x)

    # -- compute
    def codec(self):
        if self._codec is None:
            enc, dec = self._chains()
            self._codec = engine.DenseCodec(enc, dec)
        return self._codec

    def _np(self, key):
        return self._sd[key].detach().cpu().numpy().astype(np.float64)

    def encode(self, x, precision="auto"):
        z = self.codec().encode(_as_cuda_f32(x).reshape(-1, self.n_features), precision=precision)
        return z.to(self.dtype)

    def decode(self, z, precision="auto"):
        y = self.codec().decode(_as_cuda_f32(z).reshape(-1, self.z_dim), precision=precision)
        return y.to(self.dtype)

    def forward(self, x):
        return self.decode(self.encode(x))


class AE(_DenseBase):
    """reference models.py:116-183 (float64 Linear stack 24-200-100-50-z and mirror, LeakyReLU between,
    none on latent / output)."""

    names = ("en1", "en2", "en3", "en4", "de1", "de2", "de3", "de4")

    def _build(self):
        dims = (self.n_features,) + self.hidden + (self.z_dim,) + self.hidden[::-1] + (self.n_features,)
        for i, name in enumerate(self.names):
            self._linear(name, dims[i], dims[i + 1])

    def children(self):
        return [_Layer(self._sd, n) for n in self.names]

    def _chains(self):
        acts = ("leaky", "leaky", "leaky", "none")
        enc = [(self._np(n + ".weight"), self._np(n + ".bias"), a) for n, a in zip(self.names[:4], acts)]
        dec = [(self._np(n + ".weight"), self._np(n + ".bias"), a) for n, a in zip(self.names[4:], acts)]
        return enc, dec

    def linear_tensors(self):
        """(weights, biases) float64 ndarrays in layer order - what bb_trainer_create takes"""
        return [self._np(n + ".weight") for n in self.names], [self._np(n + ".bias") for n in self.names]

    def set_linear_tensors(self, weights, biases):
        for n, w, b in zip(self.names, weights, biases):
            self._sd[n + ".weight"] = torch.from_numpy(np.asarray(w)).to(self.dtype).clone()
            self._sd[n + ".bias"] = torch.from_numpy(np.asarray(b)).to(self.dtype).clone()
        self._codec = None

    # activation extraction (models.py:160-183): filled by training.train from the last forward
    def get_layers(self):
        return [_Layer(self._sd, n) for n in ("en1", "en2", "en3", "de1", "de2", "de3")]

    def store_hooks(self):
        self._hooks = True
        return ["hook%d" % i for i in range(6)]

    def get_activations(self):
        return self.activations

    def detach_hooks(self, hooks):
        self._hooks = False


class CFD_dense_AE(AE):
    """reference models.py:186-253: the float32 twin of AE."""

    dtype = torch.float32


class AE_Dropout_BN(_DenseBase):
    """reference models.py:256-313.  Eval mode: Dropout is the identity and BatchNorm1d is a per-feature
    affine map, folded here (float64) into the neighbouring Linear; the encoder keeps LeakyReLU on the
    latent, the decoder ends in BN -> ReLU."""

    enc_names = ("enc_nn.0", "enc_nn.3", "enc_nn.6", "enc_nn.9")
    dec_names = ("dec_nn.0", "dec_nn.3", "dec_nn.6", "dec_nn.9")
    bn_names = ("dec_nn.2", "dec_nn.5", "dec_nn.8", "dec_nn.10")
    dropout_p = (0.5, 0.4, 0.3, 0.2)

    def _build(self):
        dims = (self.n_features,) + self.hidden + (self.z_dim,)
        for i, name in enumerate(self.enc_names):
            self._linear(name, dims[i], dims[i + 1])
        ddims = dims[::-1]
        for i, name in enumerate(self.dec_names):
            self._linear(name, ddims[i], ddims[i + 1])
            n = ddims[i + 1]
            bn = self.bn_names[i]
            self._sd[bn + ".weight"] = torch.ones(n, dtype=self.dtype)
            self._sd[bn + ".bias"] = torch.zeros(n, dtype=self.dtype)
            self._sd[bn + ".running_mean"] = torch.zeros(n, dtype=self.dtype)
            self._sd[bn + ".running_var"] = torch.ones(n, dtype=self.dtype)
            self._sd[bn + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    def children(self):
        return ["enc_nn", "dec_nn"]

    def parameters(self):
        return [v for k, v in self._sd.items() if k.endswith((".weight", ".bias"))]

    def _chains(self):
        enc = [(self._np(n + ".weight"), self._np(n + ".bias"), "leaky") for n in self.enc_names]
        s, t = [], []
        for bn in self.bn_names:
            sc = self._np(bn + ".weight") / np.sqrt(self._np(bn + ".running_var") + BN_EPS)
            s.append(sc)
            t.append(self._np(bn + ".bias") - sc * self._np(bn + ".running_mean"))
        dec = []
        for i, n in enumerate(self.dec_names):
            w, b = self._np(n + ".weight"), self._np(n + ".bias")
            if i > 0:  # BN_{i-1} sits between LeakyReLU_{i-1} and this Linear
                b = b + w @ t[i - 1]
                w = w * s[i - 1][None, :]
            if i == 3:  # Linear -> BN -> ReLU
                w = w * s[3][:, None]
                b = b * s[3] + t[3]
            dec.append((w, b, "relu" if i == 3 else "leaky"))
        return enc, dec

    # what bb_trainer_create_dbn takes / returns
    def linear_tensors(self):
        names = self.enc_names + self.dec_names
        return [self._np(n + ".weight") for n in names], [self._np(n + ".bias") for n in names]

    def bn_tensors(self):
        out = {k: [self._np(bn + "." + k) for bn in self.bn_names] for k in ("weight", "bias", "running_mean", "running_var")}
        out["num_batches_tracked"] = [int(self._sd[bn + ".num_batches_tracked"]) for bn in self.bn_names]
        return out

    def set_linear_tensors(self, weights, biases):
        for n, w, b in zip(self.enc_names + self.dec_names, weights, biases):
            self._sd[n + ".weight"] = torch.from_numpy(np.asarray(w)).to(self.dtype).clone()
            self._sd[n + ".bias"] = torch.from_numpy(np.asarray(b)).to(self.dtype).clone()
        self._codec = None

    def set_bn_tensors(self, bn):
        for i, name in enumerate(self.bn_names):
            for k in ("weight", "bias", "running_mean", "running_var"):
                self._sd[name + "." + k] = torch.from_numpy(np.asarray(bn[k][i])).to(self.dtype).clone()
            self._sd[name + ".num_batches_tracked"] = torch.tensor(int(bn["num_batches_tracked"][i]), dtype=torch.long)
        self._codec = None

    def forward(self, x):
        if self.training:
            raise NotImplementedError("train-mode forward of AE_Dropout_BN runs inside training.train")
        return super().forward(x)

    def encode(self, x, precision="auto"):
        if self.training:
            raise NotImplementedError("train-mode forward of AE_Dropout_BN runs inside training.train")
        return super().encode(x, precision)

    def decode(self, z, precision="auto"):
        if self.training:
            raise NotImplementedError("train-mode forward of AE_Dropout_BN runs inside training.train")
        return super().decode(z, precision)
