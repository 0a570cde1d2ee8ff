"""Streaming ingest without a staging copy, the partly resident SNP-major matrix (the layout that lets 500,000 x 700,000
fit one B200), the ingest pre-selection mask, and narrowing the resident SNP set after the ingest (the EigenSNP
workflow's LD-block step, src/prepare.rs:1424-1563)."""
import numpy as np
import pytest

from oracle import pca

from helpers import make_dataset, standardized

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_in", [403, 640, 1030])       # file pitches 101, 160, 258 bytes: rows at every alignment mod 16
@pytest.mark.parametrize("chunk_rows", [None, 61])
def test_ingest_from_unaligned_rows_equals_three_call_path(gpu_ctx, monkeypatch, n_in, chunk_rows):
    """Counts and the recode read the staged rows at the FILE's pitch (no re-pitched copy): bit-identical counts,
    masks, statistics, standardized blocks and rfit against gpca_load_bed + gpca_snp_qc + gpca_set_pca_snps_mask."""
    import genomic_pca_b200 as gp
    if chunk_rows:
        monkeypatch.setenv("GPCA_INGEST_CHUNK_ROWS", str(chunk_rows))
    m = 900
    g, payload = make_dataset(n_in, m, n_pops=4, seed=5 + n_in, missing_rate=0.01)
    cfg = gp.QcConfig(0.95, 0.02, 1e-6)
    gpu_ctx.load_bed(payload, n_in, m)
    keep0, mean0, sd0, code0 = gpu_ctx.snp_qc(cfg)
    d0 = gpu_ctx.set_pca_snps_mask(keep0, mean0, sd0)
    counts0 = gpu_ctx.snp_counts()
    r0 = gpu_ctx.rfit(3, 5, 2, seed=11)
    ctx2 = gp.Context(0)
    keep1, mean1, sd1, code1, d1 = ctx2.ingest_bed(payload, n_in, m, qc=cfg)
    assert d1 == d0 and np.array_equal(keep1, keep0) and np.array_equal(code1, code0)
    assert np.array_equal(mean1, mean0) and np.array_equal(sd1, sd0)
    assert all(np.array_equal(a, b) for a, b in zip(counts0, ctx2.snp_counts()))
    assert ctx2.resident_snp_rows == d0
    r1 = ctx2.rfit(3, 5, 2, seed=11)
    assert np.array_equal(r1[1], r0[1]) and np.array_equal(r1[0], r0[0]) and np.array_equal(r1[2], r0[2])
    ctx2.close()


def _blocks(d, sizes):
    edges = [0]
    for sz in sizes:
        if edges[-1] + sz < d:
            edges.append(edges[-1] + sz)
    edges.append(d)
    return [np.arange(edges[i], edges[i + 1]) for i in range(len(edges) - 1)]


def test_partly_resident_snp_major_matrix(monkeypatch):
    """Memory budget too small for both orientations: the sample-major matrix stays whole, the SNP-major one keeps its
    first rows and re-creates the rest window by window from the other.  rfit, the standardized-block accessor and
    EigenSNP (id-order layout) must give what the fully resident layout gives: the same integer arithmetic on the same
    fields -- bit-identical for rfit (same K splits per segment are not guaranteed, but the integer accumulation is
    exact and the fp32 epilogue is per row)."""
    import genomic_pca_b200 as gp
    n, m = 1500, 9000
    g, payload = make_dataset(n, m, n_pops=5, seed=314)
    cfg = gp.QcConfig(0.9, 0.01, 1.0)
    full = gp.Context(0)
    keep, mean, sd, _, d = full.ingest_bed(payload, n, m, qc=cfg)
    assert full.resident_snp_rows == d
    r_full = full.rfit(4, 8, 2, seed=3)
    ids = np.arange(0, d, 11)
    z_full = full.get_standardized_snp_sample_block(ids)
    blocks = _blocks(d, [700, 333, 1025, 64, 2000, 511, 1500, 900])
    ecfg = gp.EigenSnpConfig(target_num_global_pcs=4, components_per_ld_block=5, subset_factor=0.5, min_subset_size=300,
                             max_subset_size=800, local_oversampling=6, global_oversampling=8, random_seed=17,
                             refine_pass_count=1)
    e_full = full.eigensnp(blocks, ecfg)
    full.close()

    # budget: the sample-major matrix + 3,584 resident rows + a ring window of 2,560 rows, no reserve
    part = gp.Context(0)
    part.set_memory_reserve(0)
    pitch_s = -(-((n + 3) // 4) // 128) * 128
    pitch_t = -(-((m + 3) // 4) // 128) * 128
    monkeypatch.setenv("GPCA_INGEST_CHUNK_ROWS", "700")
    monkeypatch.setenv("GPCA_DEBUG_WINDOW_ROWS", "2560")
    monkeypatch.setenv("GPCA_DEBUG_MEM_BUDGET", str(pitch_t * n + (3584 + 2560) * pitch_s + 64))
    keep2, mean2, sd2, _, d2 = part.ingest_bed(payload, n, m, qc=cfg)
    monkeypatch.delenv("GPCA_DEBUG_MEM_BUDGET")
    assert d2 == d and np.array_equal(keep2, keep) and np.array_equal(mean2, mean) and np.array_equal(sd2, sd)
    assert 0 < part.resident_snp_rows < d
    r_part = part.rfit(4, 8, 2, seed=3)
    assert np.abs(r_part[1] / r_full[1] - 1).max() < 1e-6
    assert pca.subspace_angle(r_part[0], r_full[0]) < 1e-5 and pca.subspace_angle(r_part[2], r_full[2]) < 1e-5
    assert np.array_equal(part.get_standardized_snp_sample_block(ids), z_full)
    e_part = part.eigensnp(blocks, ecfg)
    assert np.abs(e_part[1] / e_full[1] - 1).max() < 1e-5
    assert pca.subspace_angle(e_part[0], e_full[0]) < 1e-4 and pca.subspace_angle(e_part[2], e_full[2]) < 1e-4
    # a block list that is not made of runs needs the gathered copies, i.e. the whole SNP-major matrix
    with pytest.raises(gp.GpcaError):
        part.eigensnp([np.arange(0, d, 2), np.arange(1, d, 2)], ecfg)
    # narrowing the set afterwards is refused on a partly resident matrix (the mask goes to the ingest instead)
    with pytest.raises(gp.GpcaError):
        part.set_pca_snps(np.nonzero(keep)[0][::2], mean[keep][::2], sd[keep][::2])
    part.close()


def test_ingest_mask_and_narrowing_after_ingest(gpu_ctx):
    """SNPs outside every LD block are dropped either before the ingest (mask) or after it (gpca_set_pca_snps on the
    resident matrices, no staging copy): both must equal the three-call path on the same selection."""
    import genomic_pca_b200 as gp
    n, m = 700, 3000
    g, payload = make_dataset(n, m, n_pops=4, seed=99)
    cfg = gp.QcConfig(0.9, 0.01, 1.0)
    in_block = np.ones(m, dtype=bool)
    in_block[100:400] = False
    in_block[1234::7] = False
    # three-call reference
    gpu_ctx.load_bed(payload, n, m)
    keep, mean, sd, code = gpu_ctx.snp_qc(cfg)
    sel = keep & in_block
    idx = np.nonzero(sel)[0]
    gpu_ctx.set_pca_snps(idx, mean[idx], sd[idx])
    r0 = gpu_ctx.rfit(3, 6, 2, seed=5)
    z0 = gpu_ctx.get_standardized_snp_sample_block(np.arange(0, idx.size, 13))
    # (a) mask handed to the ingest
    a = gp.Context(0)
    a.set_ingest_mask(in_block)
    keep_a, mean_a, sd_a, code_a, d_a = a.ingest_bed(payload, n, m, qc=cfg)
    assert d_a == idx.size and np.array_equal(keep_a, sel)
    assert np.array_equal(code_a[keep & ~in_block], np.full((keep & ~in_block).sum(), 7, dtype=np.uint8))
    assert np.array_equal(code_a[~keep], code[~keep])
    ra = a.rfit(3, 6, 2, seed=5)
    assert all(np.array_equal(x, y) for x, y in zip(ra, r0))
    a.set_ingest_mask(None)
    _, _, _, _, d_all = a.ingest_bed(payload, n, m, qc=cfg)
    assert d_all == keep.sum()
    # (b) narrowing after an unmasked ingest: only the resident matrices exist
    a.set_pca_snps(idx, mean[idx], sd[idx])
    assert a.num_pca_snps == idx.size
    rb = a.rfit(3, 6, 2, seed=5)
    assert all(np.array_equal(x, y) for x, y in zip(rb, r0))
    assert np.array_equal(a.get_standardized_snp_sample_block(np.arange(0, idx.size, 13)), z0)
    # ... by mask as well, and then a SNP that is not resident any more cannot come back
    a.ingest_bed(payload, n, m, qc=cfg)
    assert a.set_pca_snps_mask(sel, mean, sd) == idx.size
    rc = a.rfit(3, 6, 2, seed=5)
    assert all(np.array_equal(x, y) for x, y in zip(rc, r0))
    with pytest.raises(gp.GpcaError):
        a.set_pca_snps(np.nonzero(keep)[0], mean[keep], sd[keep])
    a.close()


def test_ingest_bed_file_twice_on_one_context(tmp_path):
    """A context re-used for a larger file after a smaller one (the pinned read buffers grow): no stale or freed buffer
    may be read (round-1 advisor finding)."""
    import genomic_pca_b200 as gp
    ctx = gp.Context(0)
    cfg = gp.QcConfig(0.9, 0.01, 1.0)
    for n, m, seed in [(300, 500, 1), (2100, 4000, 2), (300, 500, 1)]:
        g, payload = make_dataset(n, m, n_pops=3, seed=seed)
        path = tmp_path / f"d{n}_{m}.bed"
        with open(path, "wb") as f:
            f.write(bytes([0x6c, 0x1b, 0x01]))
            f.write(payload.tobytes())
        keep, mean, sd, code, d = ctx.ingest_bed_file(str(path), n, m, qc=cfg)
        ref = gp.Context(0)
        ref.load_bed(payload, n, m)
        keep0, mean0, sd0, code0 = ref.snp_qc(cfg)
        d0 = ref.set_pca_snps_mask(keep0, mean0, sd0)
        assert d == d0 and np.array_equal(keep, keep0) and np.array_equal(mean, mean0) and np.array_equal(sd, sd0)
        r, r0 = ctx.rfit(3, 5, 2, seed=4), ref.rfit(3, 5, 2, seed=4)
        assert all(np.array_equal(x, y) for x, y in zip(r, r0))
        ref.close()
    ctx.close()
