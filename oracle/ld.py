"""LD-block file parsing and SNP -> block mapping (oracle; see oracle/__init__.py).

Follows src/prepare.rs:1424-1616 of the reference:
  * parse_ld_block_file  (:1565-1607): whitespace fields ``chr start end``; blank lines,
    lines starting with '#', 'chr\\t' or 'chromosome\\t' skipped; <3 fields skipped;
    start/end parsed as i32; tag = "{chr_norm}:{start}-{end}".
  * normalize_chromosome_name (:1610-1616): lowercase, strip every leading "chr".
  * map_snps_to_ld_blocks (:1447-1563): each QC'd SNP goes to the FIRST block (file
    order) with equal chromosome and start <= bp <= end; PcaSnpId = rank in the sorted
    set of original BIM indices; blocks sorted by tag string; ids sorted inside a block;
    blocks sharing a tag merge (HashMap keyed by tag).
"""
from __future__ import annotations

import numpy as np


def normalize_chromosome_name(name: str) -> str:
    s = name.lower()
    if s.startswith("chr"):
        while s.startswith("chr"):       # Rust trim_start_matches strips repeated prefixes
            s = s[3:]
    return s


def parse_ld_block_lines(lines):
    blocks = []
    for line in lines:
        t = line.strip()
        if (not t) or t.startswith("#") or t.startswith("chr\t") or t.startswith("chromosome\t"):
            continue
        parts = t.split()
        if len(parts) < 3:
            continue
        chrom = normalize_chromosome_name(parts[0])
        start = int(parts[1])
        end = int(parts[2])
        if not (-2**31 <= start < 2**31 and -2**31 <= end < 2**31):
            raise ValueError("LD block coordinate does not fit i32")
        blocks.append((chrom, start, end, f"{chrom}:{start}-{end}"))
    return blocks


def map_snps_to_ld_blocks(qc_original_idx, qc_chrom, qc_bp, qc_mean32, qc_sd32, parsed_blocks):
    """Inputs are per QC'd SNP in increasing original BIM index.

    Returns dict with:
      pca_original_idx  int64[D]   sorted original indices of PCA SNPs (PcaSnpId -> orig)
      mean, sd          f32[D]     in PcaSnpId order
      block_tags        list[str]  sorted
      block_snp_ids     list[np.ndarray int64]  PcaSnpIds, sorted within each block
    """
    tag_to_orig = {}
    blocked = set()
    norm = [normalize_chromosome_name(c) for c in qc_chrom]
    for i, orig in enumerate(qc_original_idx):
        c = norm[i]
        bp = int(qc_bp[i])
        for (bc, bs, be, tag) in parsed_blocks:
            if c == bc and bs <= bp <= be:
                tag_to_orig.setdefault(tag, []).append(int(orig))
                blocked.add(int(orig))
                break
    pca_orig = np.array(sorted(blocked), dtype=np.int64)
    orig_to_id = {int(o): i for i, o in enumerate(pca_orig)}
    pos_of_orig = {int(o): i for i, o in enumerate(qc_original_idx)}
    mean = np.array([qc_mean32[pos_of_orig[int(o)]] for o in pca_orig], dtype=np.float32)
    sd = np.array([qc_sd32[pos_of_orig[int(o)]] for o in pca_orig], dtype=np.float32)
    tags = sorted(t for t, v in tag_to_orig.items() if v)
    ids = [np.array(sorted(orig_to_id[o] for o in tag_to_orig[t]), dtype=np.int64) for t in tags]
    return dict(pca_original_idx=pca_orig, mean=mean, sd=sd, block_tags=tags, block_snp_ids=ids)
