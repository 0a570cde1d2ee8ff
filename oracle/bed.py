"""PLINK .bed decode, allele counts, SNP QC ladder, mean/sigma  (oracle; see oracle/__init__.py).

Follows, in the reference (`/root/reference`):
  * .bed addressing and 2-bit codes: bed-reader 1.0.6 as used at
    src/prepare.rs:622-629,682-687 with ``count_a1`` (00->2, 01->missing(-127),
    10->1, 11->0), layout restated in tests/disk.py:89-137
    (row = 3 + snp*ceil(N/4); sample i in bits 2*(i%4) of byte i//4).
  * counts / QC ladder / mean / sigma: src/prepare.rs:1216-1375.
  * HWE chi-squared p-value: src/prepare.rs:1641-1745 (statrs ChiSquared(1).cdf).
  * sample keep-list: src/prepare.rs:1058-1096.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import special

BED_MAGIC = bytes([0x6C, 0x1B, 0x01])
MISSING_I8 = -127

# code (2-bit field value) -> count_a1 dosage as i8          prepare.rs:627 (.count_a1())
_LUT_COUNT_A1 = np.array([2, MISSING_I8, 1, 0], dtype=np.int8)


def bytes_per_snp(n_samples: int) -> int:
    return (n_samples + 3) // 4


def read_bed_payload(path: str, n_samples: int, n_snps: int) -> np.ndarray:
    """Return the payload (after the 3-byte magic) as uint8 [n_snps, ceil(N/4)]."""
    raw = np.fromfile(path, dtype=np.uint8)
    if bytes(raw[:3]) != BED_MAGIC:
        raise ValueError("not a SNP-major PLINK .bed (magic 6c 1b 01)")
    bps = bytes_per_snp(n_samples)
    if raw.size != 3 + bps * n_snps:
        raise ValueError(f"bed size {raw.size} != 3 + {bps}*{n_snps}")
    return raw[3:].reshape(n_snps, bps)


def decode_codes(payload: np.ndarray, n_samples: int) -> np.ndarray:
    """uint8 [M, ceil(N/4)] -> raw 2-bit codes uint8 [M, N] (LSB-first within a byte)."""
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    m = payload.shape[0]
    out = np.empty((m, payload.shape[1] * 4), dtype=np.uint8)
    for j in range(4):
        out[:, j::4] = (payload >> (2 * j)) & 3
    return out[:, :n_samples]


def decode_count_a1(payload: np.ndarray, n_samples: int, sample_idx=None) -> np.ndarray:
    """i8 dosages [M, N] with -127 for missing, optional sample gather (iid_index)."""
    codes = decode_codes(payload, n_samples)
    if sample_idx is not None:
        codes = codes[:, np.asarray(sample_idx, dtype=np.int64)]
    return _LUT_COUNT_A1[codes]


def snp_counts(dosage_i8: np.ndarray):
    """Pass 1 of prepare.rs:1232-1279: n_valid, n(dosage 0), n(1), n(2), sum (exact)."""
    valid = dosage_i8 != MISSING_I8
    n_valid = valid.sum(axis=1).astype(np.uint32)
    n0 = (dosage_i8 == 0).sum(axis=1).astype(np.uint32)
    n1 = (dosage_i8 == 1).sum(axis=1).astype(np.uint32)
    n2 = (dosage_i8 == 2).sum(axis=1).astype(np.uint32)
    dsum = (n1.astype(np.float64) + 2.0 * n2.astype(np.float64))  # exact in f64
    return n_valid, n0, n1, n2, dsum


def chi2_1_cdf(x: float) -> float:
    """statrs 0.18 ChiSquared::new(1.0).cdf(x) = Gamma(0.5, rate 0.5).cdf(x) = P(0.5, x/2)."""
    if x <= 0.0:
        return 0.0
    if math.isinf(x):
        return 1.0
    return float(special.gammainc(0.5, 0.5 * x))


def hwe_chi_squared_p_value(hom1: int, het: int, hom2: int) -> float:
    """prepare.rs:1641-1745, statement by statement."""
    total = hom1 + het + hom2
    if total == 0:
        return 1.0
    c1 = 2.0 * float(hom1) + float(het)
    c2 = 2.0 * float(hom2) + float(het)
    tot_alleles = c1 + c2
    if tot_alleles <= 1e-9:
        return 1.0
    f1 = c1 / tot_alleles
    f2 = c2 / tot_alleles
    if f1 < 1e-9 or f2 < 1e-9:
        return 1.0
    if abs(f1 + f2 - 1.0) > 1e-6:
        return 1.0
    e1 = f1 * f1 * float(total)
    eh = 2.0 * f1 * f2 * float(total)
    e2 = f2 * f2 * float(total)
    chi = 0.0
    if e1 > 1e-9:
        chi += (float(hom1) - e1) ** 2 / e1
    elif float(hom1) > 1e-9:
        chi = math.inf
    if math.isfinite(chi):
        if eh > 1e-9:
            chi += (float(het) - eh) ** 2 / eh
        elif float(het) > 1e-9:
            chi = math.inf
    if math.isfinite(chi):
        if e2 > 1e-9:
            chi += (float(hom2) - e2) ** 2 / e2
        elif float(hom2) > 1e-9:
            chi = math.inf
    if math.isnan(chi):
        return 1.0
    if chi == math.inf:
        return 0.0
    cdf = chi2_1_cdf(chi)
    if math.isnan(cdf):
        return 1.0
    return max(1.0 - cdf, 0.0)


def _sum_sq_diff_reference_order(row_i8: np.ndarray, mean: float) -> float:
    """Pass 2 of prepare.rs:1316-1352: 32-lane chunks reduced, then added to a scalar,
    then a scalar remainder loop.  (The order *inside* one 32-lane ``reduce_sum`` is
    not specified by Rust's portable_simd; a left-to-right sum is used here.)"""
    n = row_i8.shape[0]
    acc = 0.0
    i = 0
    x = row_i8.astype(np.float64)
    valid = row_i8 != MISSING_I8
    while i + 32 <= n:
        d = x[i:i + 32] - mean
        sq = np.where(valid[i:i + 32], d * d, 0.0)
        s = 0.0
        for v in sq:
            s += float(v)
        acc += s
        i += 32
    for k in range(i, n):
        if valid[k]:
            d = float(x[k]) - mean
            acc += d * d
    return acc


def snp_qc_and_std_params(dosage_i8: np.ndarray, *, min_call_rate=0.98, min_maf=0.01,
                          max_hwe_p=1e-6, exact_order_sigma=False):
    """prepare.rs:1281-1375.  Returns (keep mask bool[M], mean f32[M], sd f32[M], fail_code u8[M]).

    fail_code: 0 kept, 1 call-rate, 2 no valid, 3 maf, 4 monomorphic, 5 hwe, 6 variance.
    mean/sd are only meaningful where keep is True (0 elsewhere).
    """
    m, n = dosage_i8.shape
    n_valid, n0, n1, n2, dsum = snp_counts(dosage_i8)
    keep = np.zeros(m, dtype=bool)
    mean32 = np.zeros(m, dtype=np.float32)
    sd32 = np.zeros(m, dtype=np.float32)
    code = np.zeros(m, dtype=np.uint8)
    for j in range(m):
        nv = int(n_valid[j])
        call_rate = float(nv) / float(n)                       # :1283
        if call_rate < min_call_rate:
            code[j] = 1
            continue
        if nv == 0:                                            # :1292
            code[j] = 2
            continue
        mean = float(dsum[j]) / float(nv)                      # :1294
        freq = mean / 2.0                                      # :1295
        maf = min(freq, 1.0 - freq)                            # :1296
        if maf < min_maf:                                      # :1299
            code[j] = 3
            continue
        if abs(freq) < 1e-9 or abs(1.0 - freq) < 1e-9:         # :1302
            code[j] = 4
            continue
        if max_hwe_p < 1.0:                                    # :1306
            p = hwe_chi_squared_p_value(int(n0[j]), int(n1[j]), int(n2[j]))
            if p <= max_hwe_p:                                 # :1310
                code[j] = 5
                continue
        if exact_order_sigma:
            ssd = _sum_sq_diff_reference_order(dosage_i8[j], mean)
        else:
            ssd = (float(n0[j]) * (0.0 - mean) ** 2 + float(n1[j]) * (1.0 - mean) ** 2
                   + float(n2[j]) * (2.0 - mean) ** 2)
        var = ssd / float(nv - 1) if nv >= 2 else 0.0          # :1357-1361
        if var <= 1e-9:                                        # :1363
            code[j] = 6
            continue
        keep[j] = True
        mean32[j] = np.float32(mean)                           # :1313
        sd32[j] = np.float32(math.sqrt(var))                   # :1364
    return keep, mean32, sd32, code


def sample_keep_indices(fam_iids, keep_file_lines=None):
    """prepare.rs:1058-1096: FAM order, membership in the keep set (exact line match)."""
    if keep_file_lines is None:
        return np.arange(len(fam_iids), dtype=np.int64)
    s = set(keep_file_lines)
    return np.array([i for i, iid in enumerate(fam_iids) if iid in s], dtype=np.int64)


def standardized_block(dosage_i8_rows: np.ndarray, mean32: np.ndarray, sd32: np.ndarray) -> np.ndarray:
    """prepare.rs:1884-2016: z = fma(x, 1/sd, -mean*(1/sd)) in f32; sd<1e-9 -> zeros;
    any missing -> error.  dosage_i8_rows is [n_snps, n_samples] already gathered."""
    if (dosage_i8_rows == MISSING_I8).any():
        raise ValueError("Unexpected missing genotype (-127i8) in SnpBlockData")
    mean32 = np.asarray(mean32, dtype=np.float32)
    sd32 = np.asarray(sd32, dtype=np.float32)
    recip = (np.float32(1.0) / sd32).astype(np.float32)               # :1948
    bterm = (-mean32 * recip).astype(np.float32)                      # :1949
    # f32 fma == round_f32(exact(x*recip + b)); exact in f64 for x in {0,1,2}
    z = (dosage_i8_rows.astype(np.float64) * recip.astype(np.float64)[:, None]
         + bterm.astype(np.float64)[:, None]).astype(np.float32)
    z[np.abs(sd32) < 1e-9, :] = 0.0                                    # :1899
    return z


def pack_codes(codes: np.ndarray) -> np.ndarray:
    """Inverse of decode_codes: uint8 codes [M, N] -> payload [M, ceil(N/4)] (pad bits 0)."""
    m, n = codes.shape
    bps = bytes_per_snp(n)
    padded = np.zeros((m, bps * 4), dtype=np.uint8)
    padded[:, :n] = codes
    out = np.zeros((m, bps), dtype=np.uint8)
    for j in range(4):
        out |= (padded[:, j::4] & 3) << (2 * j)
    return out


def dosage_to_codes(dosage: np.ndarray) -> np.ndarray:
    """count_a1 dosage (0,1,2, -127/255 missing) -> PLINK 2-bit code."""
    d = np.asarray(dosage).astype(np.int16)
    codes = np.full(d.shape, 1, dtype=np.uint8)   # missing
    codes[d == 2] = 0
    codes[d == 1] = 2
    codes[d == 0] = 3
    return codes


def qc_from_counts(n_samples: int, n_valid, n0, n1, n2, *, min_call_rate=0.98, min_maf=0.01,
                   max_hwe_p=1e-6):
    """Vectorised twin of snp_qc_and_std_params working from integer counts only
    (same f64 expressions, element-wise; used on million-SNP inputs).  Returns
    (keep, mean32, sd32, fail_code) -- identical to the scalar version (tested)."""
    n_valid = np.asarray(n_valid, dtype=np.int64)
    n0 = np.asarray(n0, dtype=np.float64)
    n1 = np.asarray(n1, dtype=np.float64)
    n2 = np.asarray(n2, dtype=np.float64)
    m = n_valid.shape[0]
    nv = n_valid.astype(np.float64)
    code = np.zeros(m, dtype=np.uint8)
    with np.errstate(all="ignore"):
        call_rate = nv / float(n_samples)
        dsum = n1 + 2.0 * n2
        mean = dsum / nv
        freq = mean / 2.0
        maf = np.minimum(freq, 1.0 - freq)
        # HWE (prepare.rs:1641-1745)
        total = n0 + n1 + n2
        c1 = 2.0 * n0 + n1
        c2 = 2.0 * n2 + n1
        tot = c1 + c2
        f1 = c1 / tot
        f2 = c2 / tot
        e1 = f1 * f1 * total
        eh = 2.0 * f1 * f2 * total
        e2 = f2 * f2 * total
        chi = np.zeros(m)
        t1 = np.where(e1 > 1e-9, (n0 - e1) ** 2 / e1, np.where(n0 > 1e-9, np.inf, 0.0))
        chi = chi + t1
        t2 = np.where(eh > 1e-9, (n1 - eh) ** 2 / eh, np.where(n1 > 1e-9, np.inf, 0.0))
        chi = np.where(np.isfinite(chi), chi + t2, chi)
        t3 = np.where(e2 > 1e-9, (n2 - e2) ** 2 / e2, np.where(n2 > 1e-9, np.inf, 0.0))
        chi = np.where(np.isfinite(chi), chi + t3, chi)
        p = np.maximum(1.0 - special.gammainc(0.5, 0.5 * np.where(np.isfinite(chi), chi, 0.0)), 0.0)
        p = np.where(np.isinf(chi), 0.0, p)
        p = np.where(np.isnan(chi), 1.0, p)
        early1 = (total == 0) | (tot <= 1e-9) | (f1 < 1e-9) | (f2 < 1e-9) | (np.abs(f1 + f2 - 1.0) > 1e-6)
        p = np.where(early1, 1.0, p)
        ssd = n0 * (0.0 - mean) ** 2 + n1 * (1.0 - mean) ** 2 + n2 * (2.0 - mean) ** 2
        var = np.where(n_valid >= 2, ssd / (nv - 1.0), 0.0)
    alive = np.ones(m, dtype=bool)

    def kill(cond, c):
        nonlocal alive
        hit = alive & cond
        code[hit] = c
        alive = alive & ~cond

    kill(call_rate < min_call_rate, 1)
    kill(n_valid == 0, 2)
    kill(maf < min_maf, 3)
    kill((np.abs(freq) < 1e-9) | (np.abs(1.0 - freq) < 1e-9), 4)
    if max_hwe_p < 1.0:
        kill(p <= max_hwe_p, 5)
    kill(var <= 1e-9, 6)
    keep = alive
    mean32 = np.where(keep, mean, 0.0).astype(np.float32)
    sd32 = np.where(keep, np.sqrt(np.where(keep, var, 0.0)), 0.0).astype(np.float32)
    return keep, mean32, sd32, code
