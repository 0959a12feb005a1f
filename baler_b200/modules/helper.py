"""Orchestration helpers with the reference's signatures (reference baler/modules/helper.py): what
`baler.py` calls for `--mode train|compress|decompress`.  Data loading and file layout are host
Python; every numeric step goes to libbaler_b200 (CUDA)."""
import argparse
import importlib
import os
import sys
from math import ceil

import numpy as np
import torch

from .. import engine
from . import data_processing, training

sys.path.append(os.getcwd())


class Config:
    """Namespace the per-project `set_config(c)` fills (reference helper.py:150-179: a dataclass used as a
    class-level namespace; attributes that are never set raise AttributeError when read).
    Extra, optional knobs of this implementation (absent = reference behaviour):
      precision        "auto" | "fp32" | "split16" | "fast"   arithmetic of the fused encode/decode kernels
      latent_dtype     "float64" (AE default, what the reference writes) | "float32" | "float16"
      l1_in_training   bool: add reg_param * L1 chain to the training loss (dead code upstream)
    """

    model_type = str  # reference helper.py:163: defaults to the *type* `str`


def get_arguments():
    """`baler --project WORKSPACE PROJECT --mode MODE [--verbose]` (reference helper.py:34-101)"""
    parser = argparse.ArgumentParser(
        prog="baler",
        description="Baler (B200-native hot path): train / compress / decompress with fused CUDA autoencoder kernels.",
        formatter_class=argparse.RawTextHelpFormatter,
    )
    parser.add_argument("--mode", type=str, required=True, help="newProject, train, compress, decompress, info")
    parser.add_argument("--project", type=str, required=True, nargs=2, metavar=("WORKSPACE", "PROJECT"),
                        help="workspace and project, e.g. --project CMS_workspace CMS_project_v1")
    parser.add_argument("--verbose", dest="verbose", action="store_true", help="Verbose mode")
    parser.set_defaults(verbose=False)
    args = parser.parse_args()
    workspace_name, project_name = args.project
    if args.mode == "newProject":
        config = None
    else:
        config = Config
        importlib.import_module(
            f"workspaces.{workspace_name}.{project_name}.config.{project_name}_config").set_config(config)
    return config, args.mode, workspace_name, project_name, args.verbose


def create_new_project(workspace_name, project_name, verbose=False, base_path="workspaces"):
    """reference helper.py:104-147: directory skeleton + default config"""
    workspace_path = os.path.join(base_path, workspace_name)
    project_path = os.path.join(workspace_path, project_name)
    if os.path.exists(project_path):
        print(f"The workspace and project ({project_path}) already exists.")
        return
    os.makedirs(project_path)
    for d in (os.path.join(workspace_path, "data"), os.path.join(project_path, "config"),
              os.path.join(project_path, "output", "compressed_output"),
              os.path.join(project_path, "output", "decompressed_output"),
              os.path.join(project_path, "output", "plotting"), os.path.join(project_path, "output", "training")):
        if verbose:
            print(f"Creating directory {d}...")
        os.makedirs(d, exist_ok=True)
    with open(os.path.join(project_path, "config", f"{project_name}_config.py"), "w") as f:
        f.write(create_default_config(workspace_name, project_name))


_DEFAULTS = [  # same keys and values as the reference's template (helper.py:182-232)
    ("input_path", None), ("data_dimension", 1), ("compression_ratio", 2.0), ("apply_normalization", True),
    ("model_name", "AE"), ("model_type", "dense"), ("epochs", 5), ("lr", 0.001), ("batch_size", 512),
    ("early_stopping", True), ("lr_scheduler", True), ("early_stopping_patience", 100), ("min_delta", 0),
    ("lr_scheduler_patience", 50), ("custom_norm", False), ("reg_param", 0.001), ("RHO", 0.05), ("test_size", 0),
    ("extra_compression", False), ("intermittent_model_saving", False), ("intermittent_saving_patience", 100),
    ("mse_avg", False), ("mse_sum", True), ("emd", False), ("l1", True), ("activation_extraction", False),
    ("deterministic_algorithm", True), ("separate_model_saving", False), ("save_error_bounded_deltas", False),
]


def create_default_config(workspace_name, project_name):
    lines = ["", "# === Configuration options ===", "", "def set_config(c):"]
    for key, val in _DEFAULTS:
        if key == "input_path":
            val = f"workspaces/{workspace_name}/data/{project_name}_data.npz"
        lines.append(f"    c.{key:<29}= {val!r}")
    return "\n".join(lines) + "\n"


def model_init(model_name):
    return data_processing.initialise_model(model_name)


def numpy_to_tensor(data):
    return torch.from_numpy(data)


def normalize(data, custom_norm):
    """reference helper.py:261-274 (per column / per pixel over axis 0)"""
    return data_processing.normalize(data, custom_norm)


def renormalize(data, true_min_list, feature_range_list):
    """reference helper.py:322-333"""
    return data_processing.renormalize_func(data, true_min_list, feature_range_list)


def process(input_path, custom_norm, test_size, apply_normalization, convert_to_blocks, verbose):
    """reference helper.py:277-319 -> (train_set, test_set, normalization_features, original_shape)"""
    data = np.load(input_path)["data"]
    if verbose:
        print("Original Dataset Shape - ", data.shape)
    original_shape = data.shape
    if convert_to_blocks:
        data = data_processing.convert_to_blocks_util(convert_to_blocks, data)
    normalization_features = data_processing.find_minmax(data)
    if apply_normalization:
        print("Normalizing the data...")
        data = normalize(data, custom_norm)
    if not test_size:
        train_set = test_set = data
    else:
        from sklearn.model_selection import train_test_split

        train_set, test_set = train_test_split(data, test_size=test_size, random_state=1)
    return train_set, test_set, normalization_features, original_shape


def train(model, number_of_columns, train_set, test_set, project_path, config):
    return training.train(model, number_of_columns, train_set, test_set, project_path, config)


def model_saver(model, model_path):
    return data_processing.save_model(model, model_path)


def detacher(tensor):
    return tensor.cpu().detach().numpy()


def get_device():
    """reference helper.py:425-439 - here a CUDA device is mandatory"""
    engine.require_cuda()
    return torch.device("cuda:0")


def _latent_np_dtype(model, config):
    name = getattr(config, "latent_dtype", None)
    if name is None:
        return np.float64 if model.dtype == torch.float64 else np.float32
    return np.dtype(name).type


def compress(model_path, config):
    """reference helper.py:473-616 -> (compressed ndarray, eb_batch, eb_deltas, eb_index).
    Normalisation uses THIS file's column min/max (helper.py:500-502), fused into the encode kernel."""
    from .. import sharded
    rank, world = sharded.dist_env()  # under torchrun: this rank's GPU, before anything touches a device
    loaded = np.load(config.input_path)
    data_before = loaded["data"]
    original_shape = data_before.shape
    if hasattr(config, "convert_to_blocks") and config.convert_to_blocks:
        data_before = data_processing.convert_to_blocks_util(config.convert_to_blocks, data_before)
    conv = False
    if config.data_dimension == 1:
        number_of_columns = len(loaded["names"])
        config.latent_space_size = ceil(number_of_columns / config.compression_ratio)
        config.number_of_columns = number_of_columns
        n_features = number_of_columns
    elif config.data_dimension == 2:
        if config.model_type == "dense":
            number_of_rows, config.number_of_columns = data_before.shape[1], data_before.shape[2]
            n_features = number_of_rows * config.number_of_columns
        else:  # convolutional (reference helper.py:521-524): sizes come from the ORIGINAL snapshot shape
            conv = True
            number_of_rows, config.number_of_columns = original_shape[1], original_shape[2]
            n_features = config.number_of_columns
        config.latent_space_size = ceil((number_of_rows * config.number_of_columns) / config.compression_ratio)
    else:
        raise NameError("Data dimension can only be 1 or 2. Got config.data_dimension = " + str(config.data_dimension))
    model = data_processing.load_model(data_processing.initialise_model(config.model_name), model_path,
                                       n_features=n_features, z_dim=config.latent_space_size)
    model.eval()
    normalise = bool(config.apply_normalization) and not config.custom_norm
    if config.apply_normalization:
        print("Normalizing...")
    table = None
    if normalise and data_before.dtype == np.float64 and data_before.size:
        # a float64 file whose offset dwarfs its spread is normalised here in float64, as the reference's numpy does
        # (data_processing.F32_OFFSET_LIMIT); the kernels then take the normalised float32 table as it is
        flat64 = data_before.reshape(data_before.shape[0], -1)
        shard_stats = None
        if world > 1:  # every rank holds the file: same decision everywhere, statistics from the row shards
            lo64, hi64 = sharded.row_range(len(flat64), rank, world)

            def shard_stats(_flat):
                f = sharded.global_minmax(_flat[lo64:hi64])
                return f[0], f[0] + f[1]
        st = data_processing.float64_stats(flat64, shard_stats)
        if st is not None:
            table = np.ascontiguousarray(data_processing.normalize_float64_host(flat64, st[0], st[1], np.float32))
            normalise = False
    if table is None:
        table = engine.host_convert(data_before.reshape(data_before.shape[0], -1), np.float32)
    codec = model.codec(data_before.shape[1], data_before.shape[2]) if conv else model.codec()
    if getattr(config, "save_error_bounded_deltas", False):
        return _compress_with_deltas(codec, model, table, normalise, config, rank, world)
    if world == 1:
        compressed, _ = codec.compress_host(table, recompute_minmax=normalise,
                                            z_dtype=_latent_np_dtype(model, config),
                                            precision=getattr(config, "precision", "auto"))
        return compressed, [], [], []
    # launched under torchrun: rows are sharded contiguously over the ranks; the only exchange on the way in is the
    # 2 x C column min / max of the whole file, rank 0 collects the latent rows (the other ranks return None)
    lo, hi = sharded.row_range(len(table), rank, world)
    shard = table[lo:hi]
    feats = sharded.global_minmax(shard) if normalise else None
    if len(shard):
        z, _ = codec.compress_host(shard, features=feats, z_dtype=_latent_np_dtype(model, config),
                                   precision=getattr(config, "precision", "auto"))
    else:
        z = np.empty((0, config.latent_space_size), dtype=_latent_np_dtype(model, config))
    return sharded.gather_rows_to_rank0(z, len(table)), [], [], []


def _compress_with_deltas(codec, model, table, normalise, config, rank=0, world=1):
    """reference helper.py:583-611 with config.save_error_bounded_deltas: every batch is encoded, decoded again and
    compared with its (normalised) input; elements whose relative error exceeds config.error_bounded_requirement percent
    get a float16 delta (helper.save_error_bounded_requirement, helper.py:442-470).  Here the table goes through the GPU in
    row chunks: encode, decode, one scan kernel (bb_error_bounded_deltas_f32); the hits are regrouped per batch of
    config.batch_size rows on the host, in the reference's (batch index, deltas, (row-in-batch, column)) lists.
    Launched under torchrun every rank scans its contiguous row range (hits carry GLOBAL row numbers), rank 0 collects the
    latent rows and the hit lists and regroups them; the other ranks return (None, [], [], [])."""
    from .. import engine, sharded
    n, bs = len(table), int(config.batch_size)
    z_dtype = _latent_np_dtype(model, config)
    lo, hi = sharded.row_range(n, rank, world)
    shard = table[lo:hi]
    compressed = np.empty((hi - lo, codec.z_dim), dtype=z_dtype)
    mn = rg = None
    if normalise and n:
        # this file's own [min; range] (helper.py:500-502), over all ranks' rows
        feats = data_processing.find_minmax(table) if world == 1 else sharded.global_minmax(shard)
        mn, rg = torch.from_numpy(np.ascontiguousarray(feats[0])).cuda(), torch.from_numpy(np.ascontiguousarray(feats[1])).cuda()
    rows, cols, deltas = [], [], []
    chunk = 1 << 20
    for r0 in range(0, hi - lo, chunk):
        x = torch.from_numpy(shard[r0:r0 + chunk]).cuda()
        z = codec.encode(x, mn, rg, precision=getattr(config, "precision", "auto"))
        y = codec.decode(z, precision=getattr(config, "precision", "auto"))  # still normalised, as upstream compares it
        r, c, d = engine.error_bounded_deltas(x, y, mn, rg, config.error_bounded_requirement, row0=lo + r0)
        rows.append(r); cols.append(c); deltas.append(d)
        compressed[r0:r0 + chunk] = z.cpu().numpy().astype(z_dtype, copy=False)
    rows = np.concatenate(rows) if rows else np.empty(0, np.int64)
    cols = np.concatenate(cols) if cols else np.empty(0, np.int64)
    deltas = np.concatenate(deltas) if deltas else np.empty(0, np.float16)
    if world > 1:
        compressed = sharded.gather_rows_to_rank0(compressed, n)
        hits = sharded.gather_objects_to_rank0((rows, cols, deltas))
        if rank != 0:
            return None, [], [], []
        rows, cols, deltas = (np.concatenate([h[i] for h in hits]) for i in range(3))  # rank order = row order
    print("Total Deltas Found - ", len(rows))
    # upstream appends an entry for EVERY batch (its `len(index) > 0` test is on a 2-tuple); batches without a hit get
    # empty lists here (upstream would reuse the previous batch's deltas or crash on the first one)
    batch_of = rows // bs
    cuts = np.searchsorted(batch_of, np.arange((n + bs - 1) // bs + 1))
    eb_batch, eb_deltas, eb_index = [], [], []
    for b in range(len(cuts) - 1):
        lo, hi = cuts[b], cuts[b + 1]
        eb_batch.append(b)
        eb_deltas.append(list(deltas[lo:hi]))
        eb_index.append((rows[lo:hi] - b * bs, cols[lo:hi]))
    return compressed, eb_batch, eb_deltas, eb_index


def _apply_deltas(decompressed, input_path_deltas, input_batch_index, batch_size, col_scale):
    """reference helper.py:655-665, 708-718: out[row][col] -= delta for every stored delta; `col_scale` = the range of each
    column when un-normalisation is fused into the decode kernel ((y - d) * range + min == y * range + min - d * range)"""
    import gzip
    loaded_deltas = np.load(gzip.GzipFile(input_path_deltas, "r"), allow_pickle=True)
    loaded_index = np.load(gzip.GzipFile(input_batch_index, "r"), allow_pickle=True)
    batches, index = loaded_index[0], loaded_index[1]
    added = 0
    for i, b in enumerate(batches):
        d = np.asarray(loaded_deltas[i], dtype=np.float64)
        r, c = np.asarray(index[i][0], dtype=np.int64), np.asarray(index[i][1], dtype=np.int64)
        if len(d) == 0:
            continue
        step = d if col_scale is None else d * np.asarray(col_scale, dtype=np.float64)[c]
        np.subtract.at(decompressed, (int(b) * batch_size + r, c), step.astype(decompressed.dtype))
        added += len(d)
    print("Total Deltas Added - ", added)
    return decompressed


def decompress(model_path, input_path, input_path_deltas, input_batch_index, model_name, config, output_path,
               original_shape, renormalize_features=None):
    """reference helper.py:619-733 -> (decompressed ndarray, names, normalization_features).
    `renormalize_features` ([min; range], optional, not in the reference signature) fuses the
    un-normalisation of baler.py:410-424 into the decode kernel."""
    from .. import sharded
    rank, world = sharded.dist_env()  # under torchrun: this rank's GPU, before anything touches a device
    loaded = np.load(input_path)
    data, names, normalization_features = loaded["data"], loaded["names"], loaded["normalization_features"]
    latent_space_size = data.shape[1]
    model_dict = torch.load(str(model_path), map_location="cpu")
    # the reference reads len() of the last state-dict entry, which crashes on AE_Dropout_BN's
    # num_batches_tracked scalar; take the last tensor that has a length instead
    number_of_columns = [len(v) for v in model_dict.values() if getattr(v, "ndim", 0) >= 1][-1]
    model = data_processing.load_model(data_processing.initialise_model(config.model_name), model_path,
                                       n_features=number_of_columns, z_dim=latent_space_size)
    model.eval()
    if data.dtype not in (np.float16, np.float32, np.float64):
        data = data.astype(np.float32)
    out_dtype = np.float64 if model.dtype == torch.float64 else np.float32
    conv = config.data_dimension == 2 and config.model_type == "convolutional"
    if conv:  # block shape: convert_to_blocks = [1, h, w], else the whole snapshot
        blocks = getattr(config, "convert_to_blocks", None)
        h, w = (int(blocks[1]), int(blocks[2])) if blocks else (original_shape[1], original_shape[2])
        codec = model.codec(h, w)
    else:
        codec = model.codec()
    host_renorm = None
    if renormalize_features is not None and out_dtype == np.float64:
        f64 = np.asarray(renormalize_features, dtype=np.float64).reshape(2, -1)
        if data_processing._ill_conditioned(f64[0], f64[0] + f64[1]):
            # float32 could not hold y * range + min (data_processing.F32_OFFSET_LIMIT): decode to normalised values and
            # un-normalise in float64 on the host, as the reference's renormalize_func does
            host_renorm, renormalize_features = f64, None
    if world == 1:
        decompressed = codec.decompress_host(data, features=renormalize_features, y_dtype=out_dtype,
                                             precision=getattr(config, "precision", "auto"))
    else:  # under torchrun: every rank decodes its contiguous row range, rank 0 collects (the others return None data)
        lo, hi = sharded.row_range(len(data), rank, world)
        part = (codec.decompress_host(np.ascontiguousarray(data[lo:hi]), features=renormalize_features, y_dtype=out_dtype,
                                      precision=getattr(config, "precision", "auto"))
                if hi > lo else np.empty((0, codec.n_features), dtype=out_dtype))
        decompressed = sharded.gather_rows_to_rank0(part, len(data))
        if decompressed is None:
            return None, names, normalization_features
    if host_renorm is not None:
        decompressed = decompressed * host_renorm[1] + host_renorm[0]
        renormalize_features = host_renorm
    if getattr(config, "save_error_bounded_deltas", False):  # host step; under torchrun rank 0 holds the gathered rows
        decompressed = _apply_deltas(decompressed, input_path_deltas, input_batch_index, int(config.batch_size),
                                     None if renormalize_features is None else renormalize_features[1])
    if conv:
        decompressed = decompressed.reshape(len(decompressed), 1, h, w)
    if config.data_dimension == 2 and config.model_type == "dense":
        decompressed = decompressed.reshape((len(decompressed), original_shape[1], original_shape[2]))
    return decompressed, names, normalization_features
