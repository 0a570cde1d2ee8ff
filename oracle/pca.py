"""Exact PCA, randomized-SVD PCA ("rfit") and EigenSNP restatements (oracle; see oracle/__init__.py).

PARITY UNPINNED for everything in this file: the arithmetic behind
``efficient_pca::PCA::{rfit,transform}`` (call sites src/main.rs:648-659) and
``efficient_pca::eigensnp::EigenSNPCoreAlgorithm::compute_pca`` (call site
src/main.rs:365, config src/main.rs:311-327) is in an external, unpinned crate whose
source is not in /root/reference.  What follows restates the published algorithms the
crate documents (Halko-Martinsson-Tropp randomized range finder with power iterations;
EigenSNP = per-LD-block local bases -> condensed features -> global rSVD -> refinement),
with the parameters the reference passes at its call sites, and is anchored on the exact
f64 eigen-decomposition below.

Layout convention used throughout: ``S`` is the standardized matrix, SNP-major,
shape [D, N] (D SNPs/variants, N samples): S[d, n] = (g[d, n] - mean[d]) / sd[d].
"""
from __future__ import annotations

import numpy as np

from . import rng

STREAM_RFIT_OMEGA = 1
STREAM_EIGENSNP_GLOBAL = 2
STREAM_EIGENSNP_SUBSET = 100
STREAM_EIGENSNP_LOCAL0 = 1000


# --------------------------------------------------------------------------- helpers
def standardize_dense(dosage, mean=None, sd=None, dtype=np.float64):
    """dosage [D, N] (no missing) -> S.  mean/sd default to per-row mean, ddof=1 sd with
    sd <= 1e-9 -> 1 (the rfit convention recalled in SURVEY.md section 8c)."""
    g = np.asarray(dosage, dtype=np.float64)
    if mean is None:
        mean = g.mean(axis=1)
    if sd is None:
        sd = g.std(axis=1, ddof=1)
        sd = np.where(sd <= 1e-9, 1.0, sd)
    mean = np.asarray(mean, dtype=np.float64)
    sd = np.asarray(sd, dtype=np.float64)
    return ((g - mean[:, None]) / sd[:, None]).astype(dtype)


def orth(y):
    """Orthonormal basis of range(y) (Householder QR)."""
    q, _ = np.linalg.qr(y)
    return q


def fix_signs(scores, *others):
    """Make the largest-|.| score of every component positive (the build's convention)."""
    scores = np.array(scores, copy=True)
    outs = [np.array(o, copy=True) for o in others]
    for j in range(scores.shape[1]):
        i = int(np.argmax(np.abs(scores[:, j])))
        if scores[i, j] < 0:
            scores[:, j] *= -1
            for o in outs:
                o[:, j] *= -1
    return (scores, *outs)


def subspace_angle(a, b):
    """Largest principal angle (rad) between column spaces of a and b."""
    qa, _ = np.linalg.qr(np.asarray(a, dtype=np.float64))
    qb, _ = np.linalg.qr(np.asarray(b, dtype=np.float64))
    s = np.linalg.svd(qa.T @ qb, compute_uv=False)
    # use sine form for small angles
    proj = qb - qa @ (qa.T @ qb)
    sn = np.linalg.svd(proj, compute_uv=False)
    return float(np.arcsin(min(1.0, sn.max()))) if s.min() > 0.5 else float(np.arccos(max(-1.0, min(1.0, s.min()))))


# --------------------------------------------------------------------------- exact
def exact_pca(S, k):
    """Exact PCA of the standardized matrix by eigh of the smaller Gram matrix.
    Returns scores [N,k] (= U*s), eigenvalues s^2/(N-1) [k], loadings [D,k] (orthonormal)."""
    S = np.asarray(S, dtype=np.float64)
    d, n = S.shape
    if n <= d:
        gram = S.T @ S
        w, v = np.linalg.eigh(gram)
        w = w[::-1][:k]
        v = v[:, ::-1][:, :k]
        s = np.sqrt(np.maximum(w, 0))
        scores = v * s
        load = (S @ v) / s
    else:
        gram = S @ S.T
        w, u = np.linalg.eigh(gram)
        w = w[::-1][:k]
        u = u[:, ::-1][:, :k]
        s = np.sqrt(np.maximum(w, 0))
        load = u
        scores = S.T @ u
    scores, load = fix_signs(scores, load)
    return scores, w / (n - 1), load


# --------------------------------------------------------------------------- rfit
def rfit(S, k, n_oversamples=10, seed=0, power_iters=2, row0=0):
    """Randomized PCA as the reference drives it (src/main.rs:636-659: k, n_oversamples=10,
    seed, tol=None; then transform on the same matrix).  S: [D, N] standardized.
    ``row0`` = global index of S's first row (SNP shard offset) for the Omega stream."""
    S = np.asarray(S, dtype=np.float64)
    d, n = S.shape
    l = min(k + n_oversamples, n, d)
    k = min(k, l)
    omega = rng.gaussian_matrix(seed, STREAM_RFIT_OMEGA, row0, d, l)
    y = S.T @ omega
    for _ in range(power_iters):
        q = orth(y)
        z = orth(S @ q)
        y = S.T @ z
    q = orth(y)
    b = S @ q                                   # [D, l]
    w, vb = np.linalg.eigh(b.T @ b)
    w = w[::-1]
    vb = vb[:, ::-1]
    s = np.sqrt(np.maximum(w, 0))
    rot = (b @ vb[:, :k]) / s[:k]               # [D, k]
    scores = S.T @ rot                          # transform(): [N, k]
    scores, rot = fix_signs(scores, rot)
    return scores, (w[:k] / (n - 1)), rot


# --------------------------------------------------------------------------- EigenSNP
def eigensnp_subset(n, seed, subset_factor=0.075, min_subset=10_000, max_subset=40_000):
    ns = int(subset_factor * n)
    ns = max(min_subset, min(ns, max_subset))
    ns = min(ns, n)
    if ns == n:
        return np.arange(n, dtype=np.int64)
    keys = rng.subset_keys(seed, STREAM_EIGENSNP_SUBSET, n)
    order = np.argsort(keys, kind="stable")
    return np.sort(order[:ns]).astype(np.int64)


def _rsvd_left_basis(x, c, oversample, q, seed, stream):
    """Top-c left singular vectors of x [m, ns] by randomized SVD."""
    m, ns = x.shape
    c = min(c, m, ns)
    l = min(c + oversample, m, ns)
    omega = rng.gaussian_matrix(seed, stream, 0, ns, l)
    y = x @ omega
    for _ in range(q):
        qq = orth(y)
        z = orth(x.T @ qq)
        y = x @ z
    qq = orth(y)
    bt = x.T @ qq                               # [ns, l]
    w, ub = np.linalg.eigh(bt.T @ bt)
    ub = ub[:, ::-1]
    return qq @ ub[:, :c]


def eigensnp(S, block_snp_ids, *, k=10, components_per_block=7, subset_factor=0.075,
             min_subset=10_000, max_subset=40_000, global_oversampling=10, global_power_iters=2,
             local_oversampling=10, local_power_iters=2, seed=2025, refine_passes=1,
             return_intermediates=False, snp_id_offset=0):
    """EigenSNP restatement (effective defaults: src/main.rs:545-588, tests/sweep_run.py:24-47).
    S: [D, N] standardized in PcaSnpId order; block_snp_ids: list of id arrays (tag-sorted)."""
    S = np.asarray(S, dtype=np.float64)
    d, n = S.shape
    sub = eigensnp_subset(n, seed, subset_factor, min_subset, max_subset)
    ssub = S[:, sub]
    feats = []
    bases = []
    keys = []
    for p, ids in enumerate(block_snp_ids):
        ids = np.asarray(ids, dtype=np.int64)
        first = int(ids[0]) + snp_id_offset      # streams are keyed by the block's first (global) PcaSnpId,
        up = _rsvd_left_basis(ssub[ids], components_per_block, local_oversampling,   # so any sharding of blocks
                              local_power_iters, seed, (STREAM_EIGENSNP_LOCAL0 + first) & 0xFFFFFFFF)  # agrees
        bases.append(up)
        feats.append(up.T @ S[ids])             # [c_p, N]
        keys.extend(first * 64 + j for j in range(up.shape[1]))
    c = np.concatenate(feats, axis=0)           # [R, N]
    cm = c.mean(axis=1, keepdims=True)
    cs = c.std(axis=1, ddof=1, keepdims=True)
    cz = np.where(cs > 1e-12, (c - cm) / np.where(cs > 1e-12, cs, 1.0), 0.0)
    r = cz.shape[0]
    lg = min(k + global_oversampling, r, n)
    kk = min(k, lg)
    omega = rng.gaussian_rows(seed, STREAM_EIGENSNP_GLOBAL, np.array(keys, dtype=np.uint64), lg)
    y = cz.T @ omega
    for _ in range(global_power_iters):
        q = orth(y)
        z = orth(cz @ q)
        y = cz.T @ z
    q = orth(y)
    b = cz @ q
    w, vb = np.linalg.eigh(b.T @ b)
    vb = vb[:, ::-1]
    v = q @ vb[:, :kk]                          # initial sample-side vectors [N, k]
    v0 = v.copy()
    if refine_passes == 0:
        lmat = S @ v
        sv = np.linalg.norm(lmat, axis=0)
        load = lmat / sv
    for _ in range(refine_passes):
        lmat = orth(S @ v)                      # [D, k]
        sc = S.T @ lmat                         # [N, k]
        w, wv = np.linalg.eigh(sc.T @ sc)
        w = w[::-1]
        wv = wv[:, ::-1]
        sv = np.sqrt(np.maximum(w, 0))
        v = (sc @ wv) / sv
        load = lmat @ wv
    scores = v * sv
    scores, load = fix_signs(scores, load)
    ev = sv ** 2 / (n - 1)
    if return_intermediates:
        return scores, ev, load, dict(subset=sub, bases=bases, condensed=cz, v0=v0)
    return scores, ev, load
