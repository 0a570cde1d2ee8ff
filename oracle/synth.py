"""Structured synthetic genotypes for parity tests (oracle-side; numpy).

Balding-Nichols populations as specified in SURVEY.md section 8(d): P equal-size
populations, ancestral allele frequency ~ U(0.05, 0.5), F_ST = 0.1, genotype ~
Binomial(2, p_pop).  Used by tests/ and the bench's CPU legs; the bench's GPU leg has its
own on-device generator with the same model.
"""
from __future__ import annotations

import numpy as np


def balding_nichols(n_samples, n_snps, n_pops=4, fst=0.1, seed=0, missing_rate=0.0):
    """Returns count_a1 dosages int8 [n_snps, n_samples] (-127 = missing) and pop labels."""
    r = np.random.default_rng(seed)
    p_anc = r.uniform(0.05, 0.5, size=n_snps)
    a = p_anc * (1 - fst) / fst
    b = (1 - p_anc) * (1 - fst) / fst
    p_pop = r.beta(a[:, None], b[:, None], size=(n_snps, n_pops))
    pops = np.arange(n_samples) * n_pops // n_samples
    p = p_pop[:, pops]
    g = (r.random((n_snps, n_samples)) < p).astype(np.int8) + (r.random((n_snps, n_samples)) < p).astype(np.int8)
    if missing_rate > 0:
        g[r.random((n_snps, n_samples)) < missing_rate] = -127
    return g, pops
