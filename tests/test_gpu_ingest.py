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


@pytest.mark.parametrize("chunk_rows", [None, 97])
@pytest.mark.parametrize("mode", ["qc", "vcf_maf"])
def test_pipelined_ingest_equals_three_call_path(gpu_ctx, monkeypatch, mode, chunk_rows):
    """gpca_ingest_bed (one streaming pass, chunked) against gpca_load_bed + QC/MAF filter + gpca_set_pca_snps_mask:
    identical masks, f32 statistics and fail codes, identical counts, identical standardized blocks and rfit."""
    import genomic_pca_b200 as gp
    if chunk_rows:
        monkeypatch.setenv("GPCA_INGEST_CHUNK_ROWS", str(chunk_rows))      # 1000 SNPs -> 11 chunks, last one short
    n_in, m = 403, 1000
    g, payload = make_dataset(n_in, m, n_pops=4, seed=77, missing_rate=0.01 if mode == "qc" else 0.0)
    keep_samples = np.setdiff1d(np.arange(n_in), [3, 77, 401]).astype(np.int64)
    cfg = gp.QcConfig(0.95, 0.02, 1e-6)
    # reference path: three calls
    gpu_ctx.load_bed(payload, n_in, m, keep_samples)
    if mode == "qc":
        keep0, mean0, sd0, code0 = gpu_ctx.snp_qc(cfg)
    else:
        keep0, mean0, sd0 = gpu_ctx.vcf_maf_filter(0.02)
        code0 = None
    d0 = gpu_ctx.set_pca_snps_mask(keep0, mean0, sd0)
    counts0 = gpu_ctx.snp_counts()
    ids = np.arange(0, d0, 7)
    blk0 = gpu_ctx.get_standardized_snp_sample_block(ids) if mode == "vcf_maf" else None
    r0 = gpu_ctx.rfit(3, 5, 2, seed=11)
    # pipelined path on a fresh context
    ctx2 = gp.Context(0)
    keep1, mean1, sd1, code1, d1 = ctx2.ingest_bed(payload, n_in, m, qc=cfg if mode == "qc" else None, vcf_maf=0.02,
                                                    keep_samples=keep_samples)
    assert d1 == d0 and np.array_equal(keep1, keep0)
    assert np.array_equal(mean1, mean0) and np.array_equal(sd1, sd0)
    if mode == "qc":
        assert np.array_equal(code1, code0)
    assert ctx2.num_samples == keep_samples.size and ctx2.num_pca_snps == d0
    counts1 = ctx2.snp_counts()
    assert all(np.array_equal(a, b) for a, b in zip(counts0, counts1))
    if blk0 is not None:
        assert np.array_equal(ctx2.get_standardized_snp_sample_block(ids), blk0)
    r1 = ctx2.rfit(3, 5, 2, seed=11)
    assert np.array_equal(r1[1], r0[1]) and np.array_equal(r1[0], r0[0]) and np.array_equal(r1[2], r0[2])
    # stats not requested
    ctx3 = gp.Context(0)
    out = ctx3.ingest_bed(payload, n_in, m, qc=cfg if mode == "qc" else None, vcf_maf=0.02, keep_samples=keep_samples,
                          want_stats=False)
    assert out[-1] == d0


def test_pipelined_ingest_errors(gpu_ctx):
    import genomic_pca_b200 as gp
    g, payload = make_dataset(50, 40, seed=1)
    with pytest.raises(gp.GpcaError):
        gpu_ctx.ingest_bed(payload, 50, 40, vcf_maf=0.6)          # nothing can pass: maf <= 0.5
    with pytest.raises(gp.GpcaError):
        gpu_ctx.ingest_bed(payload, 50, 40, keep_samples=np.array([5, 2], dtype=np.int64))


def test_synthetic_generator_is_counter_based(gpu_ctx):
    """Benchmark input generator (gpca_synth_bed_device): deterministic, shard-consistent (any row range of the whole
    matrix = the same rows generated as a shard), zero pad fields, missing rate and allele frequencies as requested."""
    import torch
    dev = torch.device("cuda", 0)
    n, m = 1003, 5000                       # n % 4 = 3: one pad field per row
    bps = (n + 3) // 4
    full = torch.empty((m, bps), dtype=torch.uint8, device=dev)
    gpu_ctx.synth_bed_device(full.data_ptr(), n, m, 0, seed=7, n_pops=5, fst=0.1, missing_rate=0.02)
    again = torch.empty_like(full)
    gpu_ctx.synth_bed_device(again.data_ptr(), n, m, 0, seed=7, n_pops=5, fst=0.1, missing_rate=0.02)
    assert torch.equal(full, again)
    shard = torch.empty((1200, bps), dtype=torch.uint8, device=dev)
    gpu_ctx.synth_bed_device(shard.data_ptr(), n, 1200, 3100, seed=7, n_pops=5, fst=0.1, missing_rate=0.02)
    assert torch.equal(shard, full[3100:4300])
    other = torch.empty_like(full)
    gpu_ctx.synth_bed_device(other.data_ptr(), n, m, 0, seed=8, n_pops=5, fst=0.1, missing_rate=0.02)
    assert not torch.equal(other, full)
    payload = full.cpu().numpy()
    assert ((payload[:, -1] >> 6) == 0).all()                           # the pad field of every row is 00
    d = bed.decode_count_a1(payload, n)                                # [m, n] int8, -127 = missing
    miss = (d == -127).mean()
    assert 0.015 < miss < 0.025
    valid = np.where(d == -127, 0, d).sum(1) / (2.0 * (d != -127).sum(1))
    assert 0.005 < valid.min() and valid.max() < 0.8 and 0.2 < valid.mean() < 0.35     # AF ~ U(0.05, 0.5) + drift
    # population structure: between-population variance of allele frequencies ~ F_ST p (1 - p)
    pops = np.arange(n) * 5 // n
    g = np.where(d == -127, np.nan, d.astype(np.float64))
    fpop = np.stack([np.nanmean(g[:, pops == k], axis=1) / 2 for k in range(5)], 1)
    p = fpop.mean(1)
    fst_hat = (fpop.var(1, ddof=1) / np.maximum(p * (1 - p), 1e-6)).mean()
    assert 0.05 < fst_hat < 0.2


def test_ingest_from_bed_file_equals_memory_ingest(gpu_ctx, tmp_path, monkeypatch):
    """gpca_ingest_bed_file streams the .bed through pinned buffers: same masks / statistics / resident state as the
    in-memory ingest; magic and size are checked."""
    import genomic_pca_b200 as gp
    monkeypatch.setenv("GPCA_INGEST_CHUNK_ROWS", "130")                    # 700 SNPs -> 6 chunks
    n, m = 257, 700
    g, payload = make_dataset(n, m, n_pops=3, seed=5, missing_rate=0.01)
    cfg = gp.QcConfig(0.95, 0.02, 1e-6)
    keep0, mean0, sd0, code0, d0 = gpu_ctx.ingest_bed(payload, n, m, qc=cfg)
    r0 = gpu_ctx.rfit(3, 5, 2, seed=3)
    path = tmp_path / "x.bed"
    path.write_bytes(bytes([0x6C, 0x1B, 0x01]) + np.ascontiguousarray(payload).tobytes())
    ctx2 = gp.Context(0)
    keep1, mean1, sd1, code1, d1 = ctx2.ingest_bed_file(str(path), n, m, qc=cfg)
    assert d1 == d0 and np.array_equal(keep1, keep0) and np.array_equal(code1, code0)
    assert np.array_equal(mean1, mean0) and np.array_equal(sd1, sd0)
    r1 = ctx2.rfit(3, 5, 2, seed=3)
    assert np.array_equal(r1[1], r0[1]) and np.array_equal(r1[0], r0[0])
    with pytest.raises(gp.GpcaError):
        ctx2.ingest_bed_file(str(path), n, m + 1, qc=cfg)                  # size does not match
    bad = tmp_path / "bad.bed"
    bad.write_bytes(bytes([0x6C, 0x1B, 0x00]) + np.ascontiguousarray(payload).tobytes())
    with pytest.raises(gp.GpcaError):
        ctx2.ingest_bed_file(str(bad), n, m, qc=cfg)                       # individual-major / wrong magic
    with pytest.raises(gp.GpcaError):
        ctx2.ingest_bed_file(str(tmp_path / "missing.bed"), n, m, qc=cfg)


def test_whole_fixture_qc_mask_equals_reference_ladder(golden_full):
    """Every variant of the reference's bundled fixture through the streaming ingest (counts on the GPU, f64 ladder on
    host threads): the keep mask must be the one the reference's own ladder (tests/pca.py:86-105) produced -- all
    1,066,557 decisions, 177,570 kept."""
    import genomic_pca_b200 as gp
    n = int(golden_full["n_samples"])
    payload = golden_full["payload"]
    ctx = gp.Context(0)
    keep, mean, sd, code, d = ctx.ingest_bed(payload, n, payload.shape[0], qc=gp.QcConfig(0.98, 0.01, 1e-6))
    assert d == 177570
    assert np.array_equal(keep, golden_full["keep_ref"])
    # and through the three-call path
    ctx.load_bed(payload, n, payload.shape[0])
    keep3, *_ = ctx.snp_qc(gp.QcConfig(0.98, 0.01, 1e-6))
    assert np.array_equal(keep3, golden_full["keep_ref"])
    ctx.close()
