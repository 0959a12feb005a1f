"""Synthetic inputs of the shapes BASELINE.json names (the real CMS blob is absent upstream:
`/root/reference/.MISSING_LARGE_BLOBS`).  Host (numpy) generators only; `bench.py` has a
device-side generator of the same marginals for the 100M-row table.

Column model of the CMS jet table (24 float32 columns; the `type_list` of
`workspaces/CMS_workspace/CMS_project_v1/config/CMS_project_v1_config.py:38-63` marks
columns 12-18 and 22-23 as integers):
  cols  0-11  lognormal(0, 1)      pt / mass like
  cols 12-18  poisson(8)           multiplicities
  cols 19-21  normal(0, 1)         eta / phi like
  cols 22-23  integers in [0, 30)
"""
import numpy as np

CMS_SEED = 20260101
CFD_SEED = 20260102
CMS_COLUMNS = 24

CMS_NAMES = np.array(
    [
        "recoPFJets_ak5PFJets__RECO.obj.%s_" % v
        for v in (
            "pt", "eta", "phi", "mass", "vx", "vy", "vz", "px", "py", "pz", "et", "energy",
            "chargedHadronMultiplicity", "neutralHadronMultiplicity", "photonMultiplicity",
            "electronMultiplicity", "muonMultiplicity", "HFHadronMultiplicity",
            "HFEMMultiplicity", "chargedEmEnergy", "chargedMuEnergy", "neutralEmEnergy",
            "chargedMultiplicity", "neutralMultiplicity",
        )
    ]
)


def cms_table(n_rows, seed=CMS_SEED):
    """n_rows x 24 float32 table with CMS-jet-like marginals."""
    rng = np.random.default_rng(seed)
    t = np.empty((n_rows, CMS_COLUMNS), dtype=np.float32)
    t[:, 0:12] = rng.lognormal(0.0, 1.0, size=(n_rows, 12))
    t[:, 12:19] = rng.poisson(8.0, size=(n_rows, 7))
    t[:, 19:22] = rng.normal(0.0, 1.0, size=(n_rows, 3))
    t[:, 22:24] = rng.integers(0, 30, size=(n_rows, 2))
    return t


def cfd_snapshots(n_snap, h=50, w=50, seed=CFD_SEED):
    """n_snap x h x w float32 flow-field-like snapshots (8 low-frequency modes + 1 % noise)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
    out = np.zeros((n_snap, h, w), dtype=np.float64)
    tt = np.arange(n_snap, dtype=np.float64)[:, None, None] / max(n_snap, 1)
    for _ in range(8):
        kx, ky = rng.integers(1, 5, size=2)
        ph, om, amp = rng.uniform(0, 2 * np.pi), rng.uniform(0.5, 4.0), rng.uniform(0.2, 1.0)
        out += amp * np.sin(2 * np.pi * (kx * xx + ky * yy)[None] + ph + 2 * np.pi * om * tt)
    out += 0.01 * rng.normal(size=out.shape)
    return out.astype(np.float32)


def cms_table_device(n_rows, seed=CMS_SEED, device="cuda", chunk=10_000_000):
    """Same marginals as cms_table, generated directly in HBM with torch's CUDA Philox generator
    (the 100M / 1B-row tables of BASELINE.json never exist on the host)."""
    import torch

    g = torch.Generator(device=device).manual_seed(seed)
    t = torch.empty((n_rows, CMS_COLUMNS), dtype=torch.float32, device=device)
    for lo in range(0, n_rows, chunk):
        hi = min(n_rows, lo + chunk)
        m = hi - lo
        t[lo:hi, 0:12] = torch.randn((m, 12), device=device, generator=g).exp_()
        t[lo:hi, 12:19] = torch.poisson(torch.full((m, 7), 8.0, device=device), generator=g)
        t[lo:hi, 19:22] = torch.randn((m, 3), device=device, generator=g)
        t[lo:hi, 22:24] = torch.randint(0, 30, (m, 2), device=device, generator=g).float()
    return t
