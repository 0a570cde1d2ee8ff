"""CPU tests: the oracle against the reference's own importable helpers / data (golden fixtures),
and internal consistency of the oracle.  No GPU."""
import json
import os

import numpy as np
import pytest

from oracle import bed, ld, pca, rng, synth, vcf

from conftest import GOLDEN


def test_decode_matches_reference_disk_py(golden_rows):
    """oracle decode == tests/disk.py decode of the reference (its convention: 00->0,10->1,11->2,01->255)."""
    n = int(golden_rows["n_samples"])
    d = bed.decode_count_a1(golden_rows["payload"], n)
    ref = golden_rows["decode_ref"].astype(np.int16)
    valid = ref != 255
    assert (d[~valid] == bed.MISSING_I8).all()
    assert (d[valid].astype(np.int16) == 2 - ref[valid]).all()      # count_a1 counts the other allele


def test_hwe_matches_reference_pca_py(golden_rows):
    """prepare.rs HWE restatement vs the reference's independent tests/pca.py::hwe_pval.
    (pca.py returns 0.0 when an expected count is exactly 0; prepare.rs returns 1.0 for monomorphic
    counts -- those degenerate triples are excluded, see prepare.rs:1665.)"""
    trip = golden_rows["hwe_triples"]
    ref = golden_rows["hwe_ref"]
    checked = 0
    for (a, b, c), r in zip(trip, ref):
        tot = 2 * (a + b + c)
        if tot == 0 or (2 * a + b) == 0 or (2 * c + b) == 0:
            continue
        p = bed.hwe_chi_squared_p_value(int(a), int(b), int(c))
        assert abs(p - r) <= 1e-12 + 1e-9 * abs(r), (a, b, c, p, r)
        checked += 1
    assert checked > 50


def test_pack_roundtrip_and_ragged():
    r = np.random.default_rng(0)
    for n in (1, 3, 4, 5, 63, 64, 65, 130):
        codes = r.integers(0, 4, size=(7, n)).astype(np.uint8)
        assert (bed.decode_codes(bed.pack_codes(codes), n) == codes).all()


def test_qc_scalar_equals_vectorized(golden_rows):
    n = int(golden_rows["n_samples"])
    d = bed.decode_count_a1(golden_rows["payload"], n)
    nv, n0, n1, n2, _ = bed.snp_counts(d)
    for hwe in (1e-6, 1.0, 0.05):
        k1, m1, s1, c1 = bed.snp_qc_and_std_params(d[:1500], max_hwe_p=hwe, exact_order_sigma=True)
        k2, m2, s2, c2 = bed.qc_from_counts(n, nv[:1500], n0[:1500], n1[:1500], n2[:1500], max_hwe_p=hwe)
        assert (k1 == k2).all() and (c1 == c2).all()
        assert (m1 == m2).all() and (s1 == s2).all()       # f32 bit-exact, reference summation order vs closed form


def test_qc_with_missing_calls():
    g, _ = synth.balding_nichols(200, 300, seed=3, missing_rate=0.03)
    k1, m1, s1, c1 = bed.snp_qc_and_std_params(g, min_call_rate=0.95, exact_order_sigma=True)
    nv, n0, n1, n2, _ = bed.snp_counts(g)
    k2, m2, s2, c2 = bed.qc_from_counts(200, nv, n0, n1, n2, min_call_rate=0.95)
    assert (k1 == k2).all() and (c1 == c2).all() and (m1 == m2).all()
    assert np.abs(s1.astype(np.float64) - s2) .max() <= 1.2e-7 * np.abs(s2).max()   # <= 1 ulp(f32)
    assert 0 < k1.sum() < 300 and (c1 == 1).any()


@pytest.mark.skipif(not os.path.exists("/root/reference/data/chr22_subset50.bed.zip"), reason="reference mount absent")
def test_whole_fixture_counts():
    """Whole chr22_subset50 fixture: code histogram and the QC survivor counts quoted in SURVEY.md."""
    import zipfile
    summ = json.load(open(os.path.join(GOLDEN, "chr22_subset50_summary.json")))
    raw = zipfile.ZipFile("/root/reference/data/chr22_subset50.bed.zip").read("chr22_subset50.bed")
    n, m = summ["n_samples"], summ["n_snps"]
    payload = np.frombuffer(raw, dtype=np.uint8, offset=3).reshape(m, summ["bytes_per_snp"])
    d = bed.decode_count_a1(payload, n)
    nv, n0, n1, n2, _ = bed.snp_counts(d)
    assert [int((d == 2).sum()), int((d == bed.MISSING_I8).sum()), int((d == 1).sum()), int((d == 0).sum())] == summ["code_hist"]
    keep, *_ = bed.qc_from_counts(n, nv, n0, n1, n2, max_hwe_p=1.0)
    assert keep.sum() == 179360
    keep, *_ = bed.qc_from_counts(n, nv, n0, n1, n2)
    assert keep.sum() == 177570


def test_whole_fixture_qc_mask_pinned_to_reference_ladder(golden_full):
    """The oracle's restatement of the Rust QC ladder (src/prepare.rs:1283-1364) against the keep mask produced by the
    reference's OWN Python ladder (tests/pca.py:86-105, executed from its source by tests/golden/make_golden.py) on
    every variant of the bundled fixture.  The two ladders agree on all 1,066,557 variants (177,570 kept): the float32
    HWE cast and the extra `maf > 1e-7` of pca.py change nothing on this data."""
    n = int(golden_full["n_samples"])
    d = bed.decode_count_a1(golden_full["payload"], n)
    nv, n0, n1, n2, _ = bed.snp_counts(d)
    keep, *_ = bed.qc_from_counts(n, nv, n0, n1, n2)
    assert golden_full["diff_idx"].size == 0
    assert int(golden_full["keep_ref"].sum()) == 177570
    assert np.array_equal(keep, golden_full["keep_ref"])


def test_standardized_block_fma_semantics():
    g, _ = synth.balding_nichols(50, 20, seed=1)
    keep, mean, sd, _ = bed.snp_qc_and_std_params(g, max_hwe_p=1.0)
    z = bed.standardized_block(g[keep], mean[keep], sd[keep])
    assert z.dtype == np.float32
    ref = (g[keep].astype(np.float64) - mean[keep].astype(np.float64)[:, None]) / sd[keep].astype(np.float64)[:, None]
    assert np.abs(z - ref).max() < 1e-6
    gm = g[keep].copy()
    gm[0, 0] = bed.MISSING_I8
    with pytest.raises(ValueError):
        bed.standardized_block(gm, mean[keep], sd[keep])


def test_ld_parse_and_map():
    lines = ["# comment", "chr\tstart\tend", "chr1 100 200", "1\t150\t400", "CHR2 1 1000 extra", "bad line", "", "chrX 5 6"]
    blocks = ld.parse_ld_block_lines(lines)
    assert [b[3] for b in blocks] == ["1:100-200", "1:150-400", "2:1-1000", "x:5-6"]
    orig = np.array([3, 5, 8, 9, 12, 20])
    chrom = ["1", "chr1", "1", "2", "3", "Chr2"]
    bp = np.array([100, 180, 300, 1000, 50, 1001])
    mean = np.arange(6, dtype=np.float32)
    sd = np.arange(6, dtype=np.float32) + 1
    r = ld.map_snps_to_ld_blocks(orig, chrom, bp, mean, sd, blocks)
    assert r["pca_original_idx"].tolist() == [3, 5, 8, 9]
    assert r["block_tags"] == ["1:100-200", "1:150-400", "2:1-1000"]
    assert [b.tolist() for b in r["block_snp_ids"]] == [[0, 1], [2], [3]]      # first match in file order wins
    assert r["mean"].tolist() == [0, 1, 2, 3]


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10."""
    assert [int(x) for x in rng.philox4x32_10(0, 0, 0, 0, 0, 0)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert [int(x) for x in rng.philox4x32_10(f, f, f, f, f, f)] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert [int(x) for x in rng.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)] == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    g = rng.gaussian_matrix(7, 1, 10, 5, 3)
    g2 = rng.gaussian_matrix(7, 1, 12, 3, 3)
    assert np.array_equal(g[2:], g2)                    # shard independence


def test_vcf_filter_semantics():
    text = "\n".join([
        "##fileformat=VCFv4.2",
        "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts1\ts2\ts3\ts4",
        "1\t10\t.\tA\tG\t.\t.\t.\tGT\t0/1\t1|1\t0/0\t0/1",          # kept
        "1\t11\t.\tAT\tG\t.\t.\t.\tGT\t0/1\t1|1\t0/0\t0/1",         # REF len 2 -> skipped
        "1\t12\t.\tA\tG,T\t.\t.\t.\tGT\t0/1\t1|1\t0/0\t0/1",        # multi-allelic -> skipped
        "1\t13\t.\tA\tG\t.\t.\t.\tGT\t0/1\t./.\t0/0\t0/1",          # missing -> dropped
        "1\t14\t.\tA\tG\t.\t.\t.\tGT\t0/0\t0/0\t0/0\t0/0",          # maf 0 -> dropped
        "1\t15\t.\tC\tT\t.\t.\t.\tGT:DP\t1/1:3\t1/1:4\t1/1:5\t0/1:6",  # maf 0.125 kept
    ])
    samples, ids, d = vcf.parse_vcf_text(text, 0.01)
    assert samples == ["s1", "s2", "s3", "s4"]
    assert ids == ["1:10:A:G", "1:15:C:T"]
    assert d.tolist() == [[1, 2, 0, 1], [2, 2, 2, 1]]
    assert vcf.maf_keep(np.array([2, 2, 2, 1]), 4, 0.125) and not vcf.maf_keep(np.array([2, 2, 2, 1]), 4, 0.126)
    assert vcf.maf_keep_from_counts(np.array([1]), np.array([3]), 4, 0.125)[0]


def test_rfit_and_eigensnp_against_exact():
    g, _ = synth.balding_nichols(500, 3000, n_pops=5, seed=2)
    keep, mean, sd, _ = bed.snp_qc_and_std_params(g, max_hwe_p=1.0)
    S = pca.standardize_dense(g[keep], mean[keep].astype(np.float64), sd[keep].astype(np.float64))
    k = 4
    sc, ev, ldg = pca.exact_pca(S, k)
    sc2, ev2, ld2 = pca.rfit(S, k, 10, seed=42, power_iters=3)
    assert np.abs(ev2 / ev - 1).max() < 1e-4
    assert pca.subspace_angle(sc, sc2) < 2e-3
    blocks = [np.arange(i, min(i + 250, S.shape[0])) for i in range(0, S.shape[0], 250)]
    sc3, ev3, ld3 = pca.eigensnp(S, blocks, k=k, min_subset=200, max_subset=300, subset_factor=0.5)
    assert np.abs(ev3 / ev - 1).max() < 5e-3
    assert pca.subspace_angle(sc, sc3) < 2e-2
    # shard independence of the rfit Omega stream
    om = rng.gaussian_matrix(42, pca.STREAM_RFIT_OMEGA, 100, 50, 14)
    assert np.array_equal(om, rng.gaussian_matrix(42, pca.STREAM_RFIT_OMEGA, 0, 150, 14)[100:])
