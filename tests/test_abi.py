"""CPU tests of the C-ABI library: it loads, exports every symbol include/gpca.h declares, and its
host-only arithmetic (HWE, LD mapping) matches the oracle.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import bed, ld

from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "gpca.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpca_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_all_symbols():
    import genomic_pca_b200 as gp
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(gp.lib, s), f"libgpca.so does not export {s}"
    assert b"sm_100a" in gp.lib.gpca_version()


def test_no_oracle_in_product():
    """The product package must never import the oracle or fall back to the CPU."""
    pkg = os.path.join(ROOT, "genomic_pca_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, fn


def test_init_without_gpu_fails_loudly():
    import genomic_pca_b200 as gp
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(gp.GpcaError):
        gp.Context(0)


def test_hwe_host_matches_oracle(golden_rows):
    import genomic_pca_b200 as gp
    trip = golden_rows["hwe_triples"]
    r = np.random.default_rng(0)
    extra = r.integers(0, 5000, size=(500, 3))
    big = r.integers(0, 400000, size=(200, 3))
    for a, b, c in np.concatenate([trip, extra, big, [[0, 0, 0], [0, 0, 5], [7, 0, 0], [0, 9, 0]]]):
        p_lib = gp.hwe_chi_squared_p_value(a, b, c)
        p_or = bed.hwe_chi_squared_p_value(int(a), int(b), int(c))
        assert abs(p_lib - p_or) <= 1e-14 + 1e-12 * abs(p_or), (a, b, c, p_lib, p_or)
        # the decision at the default threshold must be identical
        assert (p_lib <= 1e-6) == (p_or <= 1e-6)


def test_ld_mapping_host_matches_oracle():
    import genomic_pca_b200 as gp
    from genomic_pca_b200 import plink
    r = np.random.default_rng(1)
    chroms = ["1", "chr1", "Chr2", "X", "chrX", "22"]
    lines = []
    for _ in range(40):
        c = chroms[r.integers(0, len(chroms))]
        s = int(r.integers(1, 5000))
        lines.append(f"{c}\t{s}\t{s + int(r.integers(0, 800))}")
    lines += ["# x", "chr\tstart\tend", "1 10 20", "1 10 20"]     # duplicate tag merges
    parsed = ld.parse_ld_block_lines(lines)
    n = 600
    snp_chrom = [chroms[i] for i in r.integers(0, len(chroms), n)]
    bp = np.sort(r.integers(1, 6000, n)).astype(np.int32)
    orig = np.sort(r.choice(5000, n, replace=False))
    mean = r.random(n).astype(np.float32)
    sd = r.random(n).astype(np.float32)
    ref = ld.map_snps_to_ld_blocks(orig, snp_chrom, bp, mean, sd, parsed)
    norm = [plink.normalize_chromosome_name(c) for c in snp_chrom]
    pca_pos, block_of, n_pca, n_blk, order = gp.map_snps_to_ld_blocks(
        norm, bp, [b[0] for b in parsed], [b[1] for b in parsed], [b[2] for b in parsed])
    assert n_pca == len(ref["pca_original_idx"]) and n_blk == len(ref["block_tags"])
    assert orig[pca_pos >= 0].tolist() == ref["pca_original_idx"].tolist()
    tags = [f"{parsed[o][0]}:{parsed[o][1]}-{parsed[o][2]}" for o in order]
    assert tags == ref["block_tags"]
    ids = pca_pos[pca_pos >= 0]
    blk = block_of[pca_pos >= 0]
    for b in range(n_blk):
        assert ids[blk == b].tolist() == ref["block_snp_ids"][b].tolist()


def test_ld_file_parser(tmp_path):
    from genomic_pca_b200 import plink
    p = tmp_path / "ld.txt"
    p.write_text("# c\nchr\tstart\tend\nchr1 100 200\n1\t150\t400\nCHR2 1 1000 extra\nbad\n\nchrX 5 6\n")
    got = plink.parse_ld_block_file(str(p))
    exp = [(b[0], b[1], b[2]) for b in ld.parse_ld_block_lines(p.read_text().splitlines())]
    assert got == exp
