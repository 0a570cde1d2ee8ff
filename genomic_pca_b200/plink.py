"""Host-side mirror of the reference's data preparation around the hot path (Python harness).

Mirrors, with the same names and error behaviour where practical:
  MicroarrayDataPreparer::try_new / prepare_data_for_eigen_snp   src/prepare.rs:922-1056
  perform_sample_qc                                               src/prepare.rs:1058-1096
  parse_ld_block_file / normalize_chromosome_name                 src/prepare.rs:1565-1616
  output_writer::*                                                src/main.rs:682-839
The statistics themselves come from the GPU through the C ABI (counts) and libgpca's host
arithmetic (QC ladder, LD mapping); nothing here imports the oracle.
"""
from __future__ import annotations

import os

import numpy as np

from .binding import Context, QcConfig, map_snps_to_ld_blocks

BED_MAGIC = bytes([0x6C, 0x1B, 0x01])


def read_fam(path):
    iids = []
    with open(path) as f:
        for ln in f:
            p = ln.split()
            if p:
                iids.append(p[1] if len(p) > 1 else p[0])
    return iids


def read_bim(path):
    chrom, sid, bp = [], [], []
    with open(path) as f:
        for ln in f:
            p = ln.split()
            if not p:
                continue
            chrom.append(p[0])
            sid.append(p[1])
            bp.append(int(p[3]))
    return chrom, sid, np.array(bp, dtype=np.int32)


def read_bed_payload(path, n_samples, n_snps):
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size < 3 or bytes(raw[:3]) != BED_MAGIC:
        raise ValueError(f"{path}: not a SNP-major PLINK .bed")
    bps = (n_samples + 3) // 4
    if raw.size != 3 + bps * n_snps:
        raise ValueError(f"{path}: size {raw.size} != 3 + {bps} x {n_snps}")
    return raw[3:]


def normalize_chromosome_name(name: str) -> str:
    s = name.lower()
    while s.startswith("chr"):
        s = s[3:]
    return s


def parse_ld_block_file(path):
    blocks = []
    with open(path) as f:
        for line in f:
            t = line.strip()
            if (not t) or t.startswith("#") or t.startswith("chr\t") or t.startswith("chromosome\t"):
                continue
            parts = t.split()
            if len(parts) < 3:
                continue
            blocks.append((normalize_chromosome_name(parts[0]), int(parts[1]), int(parts[2])))
    return blocks


def perform_sample_qc(fam_iids, keep_file=None):
    if keep_file is None:
        return None
    with open(keep_file) as f:
        keep = set(f.read().splitlines())
    return np.array([i for i, iid in enumerate(fam_iids) if iid in keep], dtype=np.int64)


def prepare_data_for_eigen_snp(ctx: Context, payload, fam_iids, bim_chrom, bim_bp, ld_blocks, qc: QcConfig | None = None,
                               keep_samples=None):
    """load -> counts -> QC ladder -> LD mapping -> resident PCA set.  Returns a dict with the
    kept sample indices, PCA SNP original indices, mean/sd and the tag-sorted block id lists."""
    n_in, m = len(fam_iids), len(bim_chrom)
    ctx.load_bed(payload, n_in, m, keep_samples)
    keep, mean, sd, code = ctx.snp_qc(qc)
    qc_idx = np.nonzero(keep)[0]
    if qc_idx.size == 0:
        raise RuntimeError("No SNPs passed all QC filters.")
    chrom_norm = [normalize_chromosome_name(bim_chrom[i]) for i in qc_idx]
    bc = [b[0] for b in ld_blocks]
    bs = [b[1] for b in ld_blocks]
    be = [b[2] for b in ld_blocks]
    pca_pos, block_of, n_pca, n_blk, order = map_snps_to_ld_blocks(chrom_norm, np.asarray(bim_bp)[qc_idx], bc, bs, be)
    if n_pca == 0:
        raise RuntimeError("No SNPs mapped to LD blocks or all resulting blocks were empty.")
    sel = pca_pos >= 0
    pca_orig = qc_idx[sel]
    ctx.set_pca_snps(pca_orig, mean[pca_orig], sd[pca_orig])
    blk = block_of[sel]
    ids = np.arange(n_pca, dtype=np.uint64)
    block_ids = [ids[blk == b] for b in range(n_blk)]
    tags = [f"{bc[o]}:{bs[o]}-{be[o]}" for o in order]
    samples = np.arange(n_in) if keep_samples is None else np.asarray(keep_samples)
    return dict(sample_idx=samples, pca_original_idx=pca_orig, mean=mean[pca_orig], sd=sd[pca_orig],
                block_snp_ids=block_ids, block_tags=tags, fail_code=code)


# ---- output writers (src/main.rs:696-839), byte-compatible formats -----------------------------
def _fmt6(x):
    """Rust `format!("{:.6}", x)` (src/main.rs:721,755,779,832): the exact decimal expansion of the binary value rounded
    half-to-even to six places -- what Python's / C's "%.6f" print too -- except that Rust spells a NaN "NaN" (no
    sign) where C prints "nan" / "-nan".  Negative zero keeps its sign ("-0.000000"); infinities are "inf" / "-inf"."""
    if x != x:
        return "NaN"
    return f"{x:.6f}"


def write_principal_components(prefix, suffix, sample_names, scores):
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    k = scores.shape[1]
    if k == 0:
        return
    with open(f"{prefix}.{suffix}", "w") as f:
        f.write("SampleID" + "".join(f"\tPC{i}" for i in range(1, k + 1)) + "\n")
        for i, name in enumerate(sample_names):
            f.write(name + "".join("\t" + _fmt6(float(scores[i, j])) for j in range(k)) + "\n")


def write_eigenvalues(prefix, eigenvalues):
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(f"{prefix}.eigenvalues.tsv", "w") as f:
        f.write("PC\tEigenvalue\n")
        for i, v in enumerate(eigenvalues):
            f.write(f"{i + 1}\t{_fmt6(float(v))}\n")


def write_loadings(prefix, variant_ids, chroms, positions, loadings):
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    k = loadings.shape[1]
    if k == 0:
        return
    with open(f"{prefix}.eigensnp.loadings.tsv", "w") as f:
        f.write("VariantID\tChrom\tPos" + "".join(f"\tPC{i}_loading" for i in range(1, k + 1)) + "\n")
        for i in range(loadings.shape[0]):
            f.write(f"{variant_ids[i]}\t{chroms[i]}\t{int(positions[i])}"
                    + "".join("\t" + _fmt6(float(loadings[i, j])) for j in range(k)) + "\n")
