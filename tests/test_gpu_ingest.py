"""GPU parity: decode / allele counts / QC masks / mean,sd / standardized block -- bit-exact vs the oracle."""
import numpy as np
import pytest

from oracle import bed

from helpers import make_dataset, oracle_qc

pytestmark = pytest.mark.gpu


def _check_counts(ctx, g, payload, keep_samples=None):
    m, n_in = g.shape
    ctx.load_bed(payload, n_in, m, keep_samples)
    gk = g if keep_samples is None else g[:, keep_samples]
    nv, n0, n1, n2, _ = bed.snp_counts(gk)
    c = ctx.snp_counts()
    assert np.array_equal(c[0], nv) and np.array_equal(c[1], n0) and np.array_equal(c[2], n1) and np.array_equal(c[3], n2)
    return gk


def test_counts_and_qc_golden_chr22(gpu_ctx, golden_rows):
    n = int(golden_rows["n_samples"])
    payload = golden_rows["payload"]
    d = bed.decode_count_a1(payload, n)
    _check_counts(gpu_ctx, d, payload)
    for hwe in (1e-6, 1.0):
        keep, mean, sd, code = gpu_ctx.snp_qc(__import__("genomic_pca_b200").QcConfig(0.98, 0.01, hwe))
        k, mu, s, c = oracle_qc(d, max_hwe_p=hwe)
        assert np.array_equal(keep, k) and np.array_equal(code, c)
        assert np.array_equal(mean, mu) and np.array_equal(sd, s)      # f32 bit-exact


@pytest.mark.parametrize("n,m", [(1, 3), (3, 5), (5, 7), (63, 40), (64, 33), (65, 9), (257, 300), (2504, 1000), (20011, 64)])
def test_counts_ragged_shapes(gpu_ctx, n, m):
    g, payload = make_dataset(n, m, seed=n, missing_rate=0.02)
    _check_counts(gpu_ctx, g, payload)
    import genomic_pca_b200 as gp
    keep, mean, sd, code = gpu_ctx.snp_qc(gp.QcConfig(0.9, 0.01, 1e-6))
    k, mu, s, c = oracle_qc(g, min_call_rate=0.9)
    assert np.array_equal(keep, k) and np.array_equal(code, c)
    assert np.array_equal(mean, mu)
    assert np.array_equal(sd, s)


def test_counts_long_rows_block_path(gpu_ctx):
    g, payload = make_dataset(70001, 12, seed=5, missing_rate=0.01)      # pitch > 16 KiB -> CTA-per-row path
    _check_counts(gpu_ctx, g, payload)


def test_sample_keep_list(gpu_ctx):
    g, payload = make_dataset(301, 200, seed=9, missing_rate=0.01)
    keep = np.sort(np.random.default_rng(0).choice(301, 123, replace=False)).astype(np.int64)
    _check_counts(gpu_ctx, g, payload, keep)
    assert gpu_ctx.num_samples == 123


def test_vcf_variant_major_load_and_maf(gpu_ctx):
    from oracle import vcf
    g, _ = make_dataset(97, 400, seed=4)
    d = g.astype(np.uint8)
    d[5, 7] = 255                                # a missing call -> variant dropped (vcf.rs:227-242)
    gpu_ctx.load_u8_variant_major(d)
    keep, mean, sd = gpu_ctx.vcf_maf_filter(0.05)
    exp = np.array([(row <= 2).all() and vcf.maf_keep(row, 97, 0.05) for row in d])
    assert np.array_equal(keep, exp)
    gd = g[keep].astype(np.float64)
    assert np.allclose(mean[keep], gd.mean(1), rtol=1e-6)
    assert np.allclose(sd[keep], gd.std(1, ddof=1), rtol=1e-6)


def test_standardized_block_bit_exact(gpu_ctx):
    g, payload = make_dataset(130, 500, seed=11)
    gpu_ctx.load_bed(payload, 130, 500)
    keep, mean, sd, _ = gpu_ctx.snp_qc()
    idx = np.nonzero(keep)[0]
    gpu_ctx.set_pca_snps(idx, mean[idx], sd[idx])
    r = np.random.default_rng(0)
    ids = r.choice(idx.size, 77, replace=False)
    samp = r.choice(130, 50, replace=False)
    z = gpu_ctx.get_standardized_snp_sample_block(ids, samp)
    ref = bed.standardized_block(g[idx][ids][:, samp], mean[idx][ids], sd[idx][ids])
    assert np.array_equal(z, ref)
    z_all = gpu_ctx.get_standardized_snp_sample_block(np.arange(idx.size))
    assert np.array_equal(z_all, bed.standardized_block(g[idx], mean[idx], sd[idx]))


def test_standardized_block_errors_on_missing(gpu_ctx):
    import genomic_pca_b200 as gp
    g, payload = make_dataset(100, 50, seed=12, missing_rate=0.01)
    gpu_ctx.load_bed(payload, 100, 50)
    keep, mean, sd, _ = gpu_ctx.snp_qc(gp.QcConfig(0.9, 0.0, 1.0))
    idx = np.nonzero(keep)[0]
    gpu_ctx.set_pca_snps(idx, mean[idx], sd[idx])
    has_missing = np.nonzero((g[idx] == bed.MISSING_I8).any(1))[0]
    assert has_missing.size
    with pytest.raises(gp.GpcaError) as e:
        gpu_ctx.get_standardized_snp_sample_block(has_missing[:1])
    assert e.value.code == -4


def test_set_pca_snps_mask_equals_index_form(gpu_ctx):
    g, payload = make_dataset(300, 2000, seed=21, missing_rate=0.01)
    gpu_ctx.load_bed(payload, 300, 2000)
    keep, mean, sd, _ = gpu_ctx.snp_qc()
    idx = np.nonzero(keep)[0]
    gpu_ctx.set_pca_snps(idx, mean[idx], sd[idx])
    ids = np.arange(0, idx.size, 7)
    has_missing = (g[idx][ids] == bed.MISSING_I8).any(1)
    ids = ids[~has_missing]
    z1 = gpu_ctx.get_standardized_snp_sample_block(ids)
    n = gpu_ctx.set_pca_snps_mask(keep, mean, sd)
    assert n == idx.size == gpu_ctx.num_pca_snps
    z2 = gpu_ctx.get_standardized_snp_sample_block(ids)
    assert np.array_equal(z1, z2)
