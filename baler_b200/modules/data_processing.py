"""Host-side mirror of the reference's baler/modules/data_processing.py for the hot path: same
function names and argument meaning; min/max and (re)normalisation run in CUDA kernels."""
import numpy as np
import torch

from .. import engine
from . import models


def convert_to_blocks_util(blocks, data):
    """reference data_processing.py:26-34: plain reshape into (n_blocks, blocks[1], blocks[2])"""
    print("Converted Dataset to Blocks of Size - ", blocks, " from original ", data.shape)
    b1, b2 = int(blocks[1]), int(blocks[2])
    return data.reshape(int(np.prod(data.shape)) // (b1 * b2), b1, b2)


def save_model(model, model_path):
    """reference data_processing.py:37-47: torch zip-pickle of the state_dict (same keys / dtypes)"""
    torch.save(model.state_dict(), model_path)


def initialise_model(model_name):
    """reference data_processing.py:76-86"""
    return getattr(models, model_name)


def load_model(model_object, model_path, n_features, z_dim):
    """reference data_processing.py:89-110"""
    model = model_object(n_features, z_dim)
    model.load_state_dict(torch.load(str(model_path), map_location="cpu"), strict=False)
    return model


def _table(data):
    """(n, c) float32 view of 1-D / 2-D / 3-D input: statistics are per position over axis 0"""
    arr = np.asarray(data)
    flat = arr.reshape(arr.shape[0], -1) if arr.ndim > 1 else arr.reshape(-1, 1)
    return arr, np.ascontiguousarray(flat, dtype=np.float32)


def _minmax_dev(flat):
    x = torch.from_numpy(flat).cuda()
    mn, mx = engine.colminmax(x)
    return x, mn, mx


def find_minmax(data):
    """[min; max - min] per column (reference data_processing.py:113-130), computed on the GPU.
    float32 arithmetic: exact for float32 / small-integer tables (what the CMS path feeds)."""
    arr, flat = _table(data)
    _, mn, mx = _minmax_dev(flat)
    mn, mx = mn.cpu().numpy(), mx.cpu().numpy()
    out_dtype = arr.dtype if arr.dtype.kind in "iuf" else np.float32
    feats = np.array([mn, mx - mn]).astype(out_dtype)
    return feats.reshape((2,) + arr.shape[1:]) if arr.ndim > 1 else feats.reshape(2)


def normalize(data, custom_norm):
    """(x - min) / (max - min) over axis 0 (reference data_processing.py:133-153 applied per column by
    helper.normalize); identity when custom_norm."""
    arr, flat = _table(data)
    if custom_norm:
        return arr
    x, mn, mx = _minmax_dev(flat)
    out = engine.normalize_table(x, mn, mx - mn).cpu().numpy()  # max - min: one float32 subtraction per column
    out = out.reshape(arr.shape)
    return out if arr.dtype == np.float32 else out.astype(np.float64)


def renormalize_std(input_data, true_min, feature_range):
    """reference data_processing.py:171-185 (one column)"""
    return renormalize_func(np.asarray(input_data).reshape(-1, 1), [true_min], [feature_range]).reshape(-1)


def renormalize_func(norm_data, min_list, range_list):
    """y * range + min (reference data_processing.py:188-203); float32 on the GPU, returned as float64"""
    arr, flat = _table(norm_data)
    x = torch.from_numpy(flat).cuda()
    mn = torch.as_tensor(np.asarray(min_list, dtype=np.float32).reshape(-1)).cuda()
    rg = torch.as_tensor(np.asarray(range_list, dtype=np.float32).reshape(-1)).cuda()
    out = engine.normalize_table(x, mn, rg, inverse=True).cpu().numpy()
    return out.reshape(arr.shape).astype(np.float64)
