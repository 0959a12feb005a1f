"""float64 tables whose offset dwarfs their spread: the host-side float64 (re)normalisation of the reference-facing
modules (data_processing.F32_OFFSET_LIMIT) against the oracle's numpy arithmetic.  No device is touched on these paths."""
import numpy as np

from baler_b200 import synth
from baler_b200.modules import data_processing as dp
from oracle import baler_oracle as orc


def offset_table(n=5000, seed=3):
    t = synth.cms_table(n, seed=seed).astype(np.float64)
    t[:, 5] = 1.0e9 + np.arange(n) % 977  # an event counter: offset 1e9, spread ~1e3
    t[:, 11] = -4.0e7 + 0.25 * t[:, 11]
    return t


def test_well_conditioned_tables_keep_the_kernel_path():
    t = synth.cms_table(5000, seed=3)
    assert dp.float64_stats(t) is None  # float32 file
    assert dp.float64_stats(t.astype(np.float64)) is None
    assert dp.float64_stats(np.empty((0, 24))) is None


def test_offset_table_statistics_and_normalisation_match_the_oracle():
    t = offset_table()
    st = dp.float64_stats(t)
    assert st is not None
    ref = orc.find_minmax(t)
    assert np.array_equal(st[0], ref[0]) and np.array_equal(st[1], ref[1])
    feats = dp.find_minmax(t)
    assert feats.dtype == np.float64 and np.array_equal(feats, ref)
    norm = dp.normalize(t, False)
    assert norm.dtype == np.float64 and np.array_equal(norm, orc.normalize(t))
    assert dp.normalize(t, True) is not None and np.array_equal(dp.normalize(t, True), t)
    back = dp.renormalize_func(norm, feats[0], feats[1])
    assert back.dtype == np.float64 and np.array_equal(back, orc.renormalize(norm, feats[0], feats[1]))
    # what float32 arithmetic would have done to the counter column: the reason for this path
    f32 = (t[:, 5].astype(np.float32) - np.float32(ref[0][5])) / np.float32(ref[1][5])
    assert np.abs(f32 - norm[:, 5]).max() > 1e-2
    # 3-D input (snapshots): statistics per position
    t3 = t[:4992].reshape(208, 24, 24)
    f3 = dp.find_minmax(t3)
    assert f3.shape == (2, 24, 24) and np.array_equal(f3, orc.find_minmax(t3))


def test_sample_probe_never_misses_an_ill_conditioned_table():
    """every value lies inside [min, max], so a row sample with ratio <= limit - 1 proves the table's is <= limit"""
    rng = np.random.default_rng(5)
    for trial in range(200):
        n = int(rng.integers(4097, 20000))
        off = 10.0 ** rng.uniform(0, 4) * rng.choice([-1, 1])
        t = (off + rng.standard_normal((n, 3)) * 10.0 ** rng.uniform(-2, 2)).astype(np.float64)
        if trial % 3 == 0:  # a few outliers outside the sample's rows
            t[rng.integers(0, n, 5), rng.integers(0, 3, 5)] *= rng.uniform(0.5, 2.0, 5)
        mn, mx = t.min(0), t.max(0)
        truly = dp._ill_conditioned(mn, mx)
        assert (dp.float64_stats(t) is not None) == truly


def test_constant_and_nan_columns_follow_numpy():
    t = offset_table(600)
    t[:, 2] = 7.0  # 0 / 0 upstream
    norm = dp.normalize(t, False)
    with np.errstate(divide="ignore", invalid="ignore"):
        assert np.array_equal(norm, orc.normalize(t), equal_nan=True)
