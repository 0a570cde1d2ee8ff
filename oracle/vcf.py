"""VCF genotype -> dosage and the MAF filter (oracle; see oracle/__init__.py).

Follows src/vcf.rs of the reference: biallelic single-base REF/ALT only (:109-121);
GT string -> 0/1/2 (:52-63; '/' or '|' separator, alleles '0'/'1' only); a variant with
any missing/unparsable GT is dropped (:172-176,201-207,227-242); MAF filter
p = sum/(2N) in f64, maf=min(p,1-p), keep iff maf >= threshold (default 0.01) (:244-266);
id = "chr:pos:ref:alt" (:268-273); matrix order = file order, files in sorted-path
order (src/main.rs:152, vcf.rs:293-315).  The text parser here is a minimal stand-in for
noodles-vcf (plain-text VCF only) -- the *statistics* are what the GPU path must match.
"""
from __future__ import annotations

import numpy as np


def parse_gt(gt: str):
    b = gt.encode()
    if len(b) != 3 or b[1] not in (0x2F, 0x7C):
        return None
    if b[0] not in (0x30, 0x31) or b[2] not in (0x30, 0x31):
        return None
    return (b[0] - 0x30) + (b[2] - 0x30)


def maf_keep(dosages_u8: np.ndarray, n_samples: int, maf_threshold: float = 0.01) -> bool:
    allele_sum = int(np.asarray(dosages_u8, dtype=np.uint32).sum())
    total = 2 * n_samples
    if total == 0:
        return False
    p = float(allele_sum) / float(total)
    return min(p, 1.0 - p) >= maf_threshold


def maf_keep_from_counts(n1, n2, n_samples: int, maf_threshold: float = 0.01):
    """Vectorised: the same expression from integer counts."""
    s = np.asarray(n1, dtype=np.float64) + 2.0 * np.asarray(n2, dtype=np.float64)
    p = s / float(2 * n_samples)
    return np.minimum(p, 1.0 - p) >= maf_threshold


def parse_vcf_text(text: str, maf_threshold: float = 0.01):
    """Returns (sample_names, variant_ids, dosages u8 [D, N]) after the reference's filters."""
    samples = None
    ids, rows = [], []
    for line in text.splitlines():
        if line.startswith("##") or not line:
            continue
        f = line.split("\t")
        if line.startswith("#CHROM"):
            samples = f[9:]
            continue
        chrom, pos, _id, ref, alt = f[0], f[1], f[2], f[3], f[4]
        if len(ref) != 1 or "," in alt or alt == ".":
            continue
        fmt = f[8].split(":")
        if "GT" not in fmt:
            continue
        gi = fmt.index("GT")
        g = []
        ok = True
        for s in f[9:]:
            parts = s.split(":")
            v = parse_gt(parts[gi]) if gi < len(parts) else None
            if v is None:
                ok = False
                break
            g.append(v)
        if not ok or len(g) != len(samples):
            continue
        g = np.array(g, dtype=np.uint8)
        if not maf_keep(g, len(samples), maf_threshold):
            continue
        ids.append(f"{chrom}:{pos}:{ref}:{alt}")
        rows.append(g)
    d = np.stack(rows) if rows else np.zeros((0, len(samples or [])), dtype=np.uint8)
    return samples, ids, d
