"""Shared builders for the parity tests (test-side only; may import oracle/)."""
import numpy as np

from oracle import bed as obed
from oracle import pca as opca
from oracle import synth


def make_dataset(n_samples, n_snps, n_pops=4, seed=0, missing_rate=0.0):
    """Synthetic structured genotypes -> (dosage i8 [M,N], PLINK payload u8 [M, ceil(N/4)])."""
    g, pops = synth.balding_nichols(n_samples, n_snps, n_pops=n_pops, seed=seed, missing_rate=missing_rate)
    payload = obed.pack_codes(obed.dosage_to_codes(g))
    return g, payload


def oracle_qc(g, **kw):
    nv, n0, n1, n2, _ = obed.snp_counts(g)
    return obed.qc_from_counts(g.shape[1], nv, n0, n1, n2, **kw)


def standardized(g_rows, mean32, sd32):
    """f64 standardized matrix with missing -> 0 (mean imputation), from f32 mean/sd as the GPU uses."""
    x = g_rows.astype(np.float64)
    miss = g_rows == obed.MISSING_I8
    inv = np.where(np.abs(sd32) < 1e-9, 0.0, 1.0 / sd32.astype(np.float64))
    s = (x - mean32.astype(np.float64)[:, None]) * inv[:, None]
    s[miss] = 0.0
    return s
