"""Counter-based Gaussian test matrices (oracle twin of csrc/philox.cuh).

The reference draws its test matrices from a seeded ChaCha stream inside the
external ``efficient_pca`` crate (call sites src/main.rs:637,648-656 ``seed`` and
src/main.rs:321 ``random_seed``); that stream cannot be reproduced here (source
absent, parity unpinned), so the build defines its own: Philox4x32-10 keyed by the
seed, counter = (row_lo, row_hi, col, stream), one standard normal per counter by
Box-Muller.  Any shard of any GPU regenerates its rows independently.
"""
from __future__ import annotations

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32-valued arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & _MASK
    c1 = np.asarray(c1, dtype=np.uint64) & _MASK
    c2 = np.asarray(c2, dtype=np.uint64) & _MASK
    c3 = np.asarray(c3, dtype=np.uint64) & _MASK
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0 = p0 >> np.uint64(32)
        lo0 = p0 & _MASK
        hi1 = p1 >> np.uint64(32)
        lo1 = p1 & _MASK
        n0 = (hi1 ^ c1 ^ np.uint64(k0)) & _MASK
        n1 = lo1
        n2 = (hi0 ^ c3 ^ np.uint64(k1)) & _MASK
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def _normals_from_keys(seed: int, stream: int, rows_u64: np.ndarray, n_cols: int) -> np.ndarray:
    """rows_u64 [R,1] uint64 row keys -> f64 [R, n_cols].  One Philox call per (row, group of 4 columns):
    columns 4g, 4g+1 = sqrt(-2 ln u1a) * (cos, sin)(2 pi u2a), columns 4g+2, 4g+3 = the same from the second pair."""
    ng = (n_cols + 3) // 4
    groups = np.arange(ng, dtype=np.uint64)[None, :]
    x0, x1, x2, x3 = philox4x32_10(rows_u64 & _MASK, rows_u64 >> np.uint64(32), groups, np.uint64(stream),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)

    def pair(a, b):
        u1 = ((a >> np.uint32(8)).astype(np.float64) + 0.5) * (1.0 / 16777216.0)
        u2 = (b >> np.uint32(8)).astype(np.float64) * (1.0 / 16777216.0)
        rad = np.sqrt(-2.0 * np.log(u1))
        return rad * np.cos(2.0 * np.pi * u2), rad * np.sin(2.0 * np.pi * u2)

    c0, c1 = pair(x0, x1)
    c2, c3 = pair(x2, x3)
    out = np.stack([c0, c1, c2, c3], axis=2).reshape(rows_u64.shape[0], ng * 4)
    return out[:, :n_cols]


def gaussian_matrix(seed: int, stream: int, row0: int, n_rows: int, n_cols: int) -> np.ndarray:
    """f64 [n_rows, n_cols]; element (r, c) depends only on (seed, stream, row0+r, c)."""
    rows = (np.arange(n_rows, dtype=np.uint64) + np.uint64(row0))[:, None]
    return _normals_from_keys(seed, stream, rows, n_cols)


def gaussian_rows(seed: int, stream: int, row_keys: np.ndarray, n_cols: int) -> np.ndarray:
    """Same generator with an explicit 64-bit key per row (EigenSNP condensed-feature test matrix)."""
    rows = np.asarray(row_keys, dtype=np.uint64)[:, None]
    return _normals_from_keys(seed, stream, rows, n_cols)


def subset_keys(seed: int, stream: int, n: int) -> np.ndarray:
    """uint32 key per sample used to pick the EigenSNP local-basis subset."""
    idx = np.arange(n, dtype=np.uint64)
    x0, _, _, _ = philox4x32_10(idx & _MASK, idx >> np.uint64(32), np.uint64(0), np.uint64(stream),
                                seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return x0
