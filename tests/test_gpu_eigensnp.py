"""GPU parity of the EigenSNP driver against the oracle restatement (same Philox streams, same subset),
and of the whole BED -> QC -> LD map -> EigenSNP flow on the golden chr22 rows."""
import numpy as np
import pytest

from oracle import bed, ld, pca

from helpers import make_dataset, standardized

pytestmark = pytest.mark.gpu


def _prep(ctx, n, m, pops, seed, missing=0.0):
    import genomic_pca_b200 as gp
    g, payload = make_dataset(n, m, n_pops=pops, seed=seed, missing_rate=missing)
    ctx.load_bed(payload, n, m)
    keep, mean, sd, _ = ctx.snp_qc(gp.QcConfig(0.9, 0.01, 1.0))
    idx = np.nonzero(keep)[0]
    ctx.set_pca_snps(idx, mean[idx], sd[idx])
    return standardized(g[idx], mean[idx], sd[idx])


@pytest.mark.parametrize("engine", [0, 1, 2])
def test_eigensnp_matches_oracle(gpu_ctx, engine):
    import genomic_pca_b200 as gp
    S = _prep(gpu_ctx, 1200, 5000, 5, seed=21)
    d = S.shape[0]
    gpu_ctx.set_sketch_engine(engine)
    # ragged blocks, not aligned to anything, last one short
    edges = list(range(0, d, 333)) + [d]
    blocks = [np.arange(edges[i], edges[i + 1]) for i in range(len(edges) - 1)]
    cfg = gp.EigenSnpConfig(target_num_global_pcs=4, components_per_ld_block=5, subset_factor=0.5, min_subset_size=300,
                            max_subset_size=700, local_oversampling=6, global_oversampling=8, random_seed=77,
                            refine_pass_count=2)
    sc, ev, load = gpu_ctx.eigensnp(blocks, cfg)
    sc_o, ev_o, ld_o = pca.eigensnp(S, blocks, k=4, components_per_block=5, subset_factor=0.5, min_subset=300,
                                    max_subset=700, local_oversampling=6, global_oversampling=8, seed=77,
                                    refine_passes=2)
    assert sc.shape == (1200, 4) and load.shape == (d, 4)
    assert np.abs(ev / ev_o - 1).max() < 1e-4
    assert pca.subspace_angle(sc, sc_o) < 1e-3
    assert pca.subspace_angle(load, ld_o) < 1e-3
    assert np.abs(sc - sc_o).max() / np.abs(sc_o).max() < 5e-3
    # and it is a good PCA: close to the exact decomposition
    sc_x, ev_x, ld_x = pca.exact_pca(S, 4)
    assert np.abs(ev / ev_x - 1).max() < 2e-3


def test_eigensnp_batched_blocks_equal_per_block_path(gpu_ctx):
    """All LD blocks in one launch per stage (item mode of the integer engine) against the one-block-at-a-time path and
    the oracle; block sizes from 1 SNP to > 2 row groups, none aligned to the 256-field stage."""
    import genomic_pca_b200 as gp
    S = _prep(gpu_ctx, 900, 4000, 5, seed=33)
    d = S.shape[0]
    sizes = [3, 40, 256, 300, 700, 64, 1, 513, 129]
    edges = [0]
    for sz in sizes:
        if edges[-1] + sz < d:
            edges.append(edges[-1] + sz)
    edges.append(d)
    blocks = [np.arange(edges[i], edges[i + 1]) for i in range(len(edges) - 1)]
    cfg = gp.EigenSnpConfig(target_num_global_pcs=4, components_per_ld_block=5, subset_factor=0.5, min_subset_size=300,
                            max_subset_size=700, local_oversampling=6, global_oversampling=8, random_seed=9,
                            refine_pass_count=1)
    gpu_ctx.set_sketch_engine(2)
    gpu_ctx.set_batch_blocks(True)
    l0 = gpu_ctx.launch_count
    sc_b, ev_b, ld_b = gpu_ctx.eigensnp(blocks, cfg)
    n_batched = gpu_ctx.launch_count - l0
    gpu_ctx.set_batch_blocks(False)
    l0 = gpu_ctx.launch_count
    sc_s, ev_s, ld_s = gpu_ctx.eigensnp(blocks, cfg)
    n_single = gpu_ctx.launch_count - l0
    gpu_ctx.set_batch_blocks(True)
    assert n_batched < n_single / 3          # the per-block loop is gone
    assert np.abs(ev_b / ev_s - 1).max() < 1e-5
    assert pca.subspace_angle(sc_b, sc_s) < 1e-4
    assert pca.subspace_angle(ld_b, ld_s) < 1e-4
    sc_o, ev_o, ld_o = pca.eigensnp(S, blocks, k=4, components_per_block=5, subset_factor=0.5, min_subset=300,
                                    max_subset=700, local_oversampling=6, global_oversampling=8, seed=9,
                                    refine_passes=1)
    assert np.abs(ev_b / ev_o - 1).max() < 1e-4
    assert pca.subspace_angle(sc_b, sc_o) < 1e-3
    # (blocks of 1 and 3 SNPs keep their whole span: their condensed rows are any rotation of it, and the row
    #  standardisation that follows is not rotation invariant -> the loadings are a little looser here than in
    #  test_eigensnp_matches_oracle; the two device paths above agree to 1e-4)
    assert pca.subspace_angle(ld_b, ld_o) < 4e-3


def test_eigensnp_unordered_blocks_and_refine0(gpu_ctx):
    """Blocks given in tag-sorted (not genomic) order with interleaved ids; refine_pass_count = 0 path."""
    import genomic_pca_b200 as gp
    S = _prep(gpu_ctx, 500, 1500, 4, seed=5)
    d = S.shape[0]
    ids = np.arange(d)
    blocks = [ids[2::3], ids[0::3], ids[1::3]]          # interleaved, each sorted
    # components_per_block = 3 = the number of structured local components (4 populations): every retained local
    # singular vector is then well separated from the noise bulk, so the condensed rows are well determined
    cfg = gp.EigenSnpConfig(target_num_global_pcs=3, components_per_ld_block=3, subset_factor=1.0, min_subset_size=10,
                            max_subset_size=100000, random_seed=3, refine_pass_count=0)
    sc, ev, load = gpu_ctx.eigensnp(blocks, cfg)
    sc_o, ev_o, ld_o = pca.eigensnp(S, blocks, k=3, components_per_block=3, subset_factor=1.0, min_subset=10,
                                    max_subset=100000, seed=3, refine_passes=0)
    # without a refinement pass there is no final Rayleigh-Ritz step: the three near-degenerate components
    # (4 equal populations) may rotate inside their subspace, so compare rotation-invariant quantities
    assert abs(ev.sum() / ev_o.sum() - 1) < 1e-3
    assert pca.subspace_angle(sc, sc_o) < 2e-3
    assert pca.subspace_angle(load, ld_o) < 2e-3


def test_eigensnp_argument_errors(gpu_ctx):
    import genomic_pca_b200 as gp
    S = _prep(gpu_ctx, 100, 300, 3, seed=6)
    d = S.shape[0]
    with pytest.raises(gp.GpcaError):
        gpu_ctx.eigensnp([], gp.EigenSnpConfig())
    with pytest.raises(gp.GpcaError):
        gpu_ctx.eigensnp([np.array([0, 1, d + 5])], gp.EigenSnpConfig())
    with pytest.raises(gp.GpcaError):
        gpu_ctx.eigensnp([np.array([0, 1]), np.array([1, 2])], gp.EigenSnpConfig())      # id listed twice


def test_bed_to_eigensnp_flow_on_chr22_rows(gpu_ctx, golden_rows, tmp_path):
    """BASELINE config 2 in miniature: chr22_subset50 rows (64 samples), synthesized BIM, one LD block,
    effective CLI defaults except components_per_block >= k.  The fixture's spectrum is flat, so only
    oracle-vs-GPU agreement is asserted (same Omega), not closeness to the exact PCA."""
    import genomic_pca_b200 as gp
    from genomic_pca_b200 import plink
    n = int(golden_rows["n_samples"])
    payload = golden_rows["payload"][:4096]
    m = payload.shape[0]
    iids = [str(x) for x in golden_rows["iids"]]
    chrom = ["22"] * m
    bp = np.arange(1, m + 1, dtype=np.int32)
    ldfile = tmp_path / "blocks.txt"
    ldfile.write_text("chr22 1 2000000000\n")
    blocks = plink.parse_ld_block_file(str(ldfile))
    prep = plink.prepare_data_for_eigen_snp(gpu_ctx, payload, iids, chrom, bp, blocks)
    # oracle side of the same preparation
    dos = bed.decode_count_a1(payload, n)
    nv, n0, n1, n2, _ = bed.snp_counts(dos)
    keep, mean, sd, _ = bed.qc_from_counts(n, nv, n0, n1, n2)
    qidx = np.nonzero(keep)[0]
    ref = ld.map_snps_to_ld_blocks(qidx, [chrom[i] for i in qidx], bp[qidx], mean[qidx], sd[qidx],
                                   ld.parse_ld_block_lines(["chr22 1 2000000000"]))
    assert np.array_equal(prep["pca_original_idx"], ref["pca_original_idx"])
    assert np.array_equal(prep["mean"], ref["mean"]) and np.array_equal(prep["sd"], ref["sd"])
    assert prep["block_tags"] == ref["block_tags"] == ["22:1-2000000000"]
    assert [b.tolist() for b in prep["block_snp_ids"]] == [b.tolist() for b in ref["block_snp_ids"]]
    S = standardized(dos[qidx], mean[qidx], sd[qidx])
    cfg = gp.EigenSnpConfig(target_num_global_pcs=10, components_per_ld_block=12)
    sc, ev, load = gpu_ctx.eigensnp(prep["block_snp_ids"], cfg)
    sc_o, ev_o, ld_o = pca.eigensnp(S, ref["block_snp_ids"], k=10, components_per_block=12)
    assert sc.shape == (64, 10)
    # This fixture is ill-conditioned for any randomized method (flat spectrum, N = 64, one block), so
    # GPU-vs-oracle agreement is only loose; what must hold exactly are the Rayleigh-Ritz identities of the final
    # refinement pass, checked here against the dense f64 standardized matrix:
    load64 = load.astype(np.float64)
    assert np.abs(load64.T @ load64 - np.eye(10)).max() < 2e-3                      # orthonormal loadings
    proj = S.T @ load64                                                             # scores = S^T loadings
    assert np.abs(proj - sc).max() / np.abs(proj).max() < 3e-3
    assert np.abs((sc.astype(np.float64) ** 2).sum(0) / (n - 1) / ev - 1).max() < 2e-3   # eigenvalues = |scores|^2/(N-1)
    g = sc.astype(np.float64).T @ sc.astype(np.float64)
    assert np.abs(g - np.diag(np.diag(g))).max() / np.diag(g).max() < 2e-3          # orthogonal score columns
    ev_x = pca.exact_pca(S, 10)[1]
    assert (ev <= ev_x * (1 + 1e-3)).all()                                          # Cauchy interlacing vs the exact PCA
    assert abs(ev.sum() / ev_o.sum() - 1) < 0.15                                    # loose: same algorithm, same seeds
    # outputs in the reference's formats
    plink.write_principal_components(str(tmp_path / "out"), "eigensnp.pca.tsv", iids, sc)
    plink.write_eigenvalues(str(tmp_path / "out"), ev)
    lines = (tmp_path / "out.eigensnp.pca.tsv").read_text().splitlines()
    assert lines[0].split("\t")[:3] == ["SampleID", "PC1", "PC2"] and len(lines) == 65
    assert (tmp_path / "out.eigenvalues.tsv").read_text().splitlines()[0] == "PC\tEigenvalue"


def test_eigensnp_shifted_copy_and_block_groups(gpu_ctx, monkeypatch):
    """Blocks that are runs of consecutive SNP ids take the shifted-copy layout path and the grouped condensed-feature
    pass.  The shifted copy must reproduce the gather + transpose layout exactly (same results bit for bit); grouping
    changes the quantisation scale of the block-diagonal operand (one per group instead of one per block), so it is
    compared within the parity tolerances.  Block sizes straddle the 64-field slot padding and the 256-field stage; the
    component counts make groups of 1 to 4 blocks."""
    import genomic_pca_b200 as gp
    S = _prep(gpu_ctx, 1100, 4200, 5, seed=41)
    d = S.shape[0]
    sizes = [65, 63, 64, 1, 255, 257, 300, 130, 5, 700, 412, 412, 412, 412, 100]
    edges = [0]
    for sz in sizes:
        if edges[-1] + sz < d:
            edges.append(edges[-1] + sz)
    edges.append(d)
    blocks = [np.arange(edges[i], edges[i + 1]) for i in range(len(edges) - 1)]
    cfg = gp.EigenSnpConfig(target_num_global_pcs=4, components_per_ld_block=7, subset_factor=0.5, min_subset_size=300,
                            max_subset_size=700, local_oversampling=6, global_oversampling=8, random_seed=13,
                            refine_pass_count=1)
    gpu_ctx.set_sketch_engine(2)
    gpu_ctx.set_batch_blocks(True)
    for v in ("GPCA_DEBUG_NO_SHIFT_COPY", "GPCA_DEBUG_NO_GROUPS", "GPCA_DEBUG_NO_ID_ORDER"):
        monkeypatch.delenv(v, raising=False)
    sc, ev, load = gpu_ctx.eigensnp(blocks, cfg)
    # run to run: bit for bit (this configuration caught a wide row store of the sketch epilogue spilling zeros into the
    # neighbouring blocks' columns of the condensed matrix -- a race between work items)
    sc_r, ev_r, load_r = gpu_ctx.eigensnp(blocks, cfg)
    assert np.array_equal(sc, sc_r) and np.array_equal(ev, ev_r) and np.array_equal(load, load_r)
    # refinement on the slot-ordered copies instead of the resident matrices in PcaSnpId order: same arithmetic in a
    # different row order (only the f64 Gram partial sums are grouped differently)
    monkeypatch.setenv("GPCA_DEBUG_NO_ID_ORDER", "1")
    sc_s, ev_s, load_s = gpu_ctx.eigensnp(blocks, cfg)
    assert np.abs(ev / ev_s - 1).max() < 1e-6
    assert pca.subspace_angle(sc, sc_s) < 1e-5 and pca.subspace_angle(load, load_s) < 1e-5
    # the shifted copy must reproduce the gather + transpose layout exactly
    monkeypatch.setenv("GPCA_DEBUG_NO_SHIFT_COPY", "1")
    sc_t, ev_t, load_t = gpu_ctx.eigensnp(blocks, cfg)
    assert np.array_equal(sc_s, sc_t) and np.array_equal(ev_s, ev_t) and np.array_equal(load_s, load_t)
    monkeypatch.setenv("GPCA_DEBUG_NO_GROUPS", "1")
    sc_g, ev_g, load_g = gpu_ctx.eigensnp(blocks, cfg)
    assert np.abs(ev / ev_g - 1).max() < 1e-5
    assert pca.subspace_angle(sc, sc_g) < 1e-4
    assert pca.subspace_angle(load, load_g) < 1e-4
    sc_o, ev_o, ld_o = pca.eigensnp(S, blocks, k=4, components_per_block=7, subset_factor=0.5, min_subset=300,
                                    max_subset=700, local_oversampling=6, global_oversampling=8, seed=13,
                                    refine_passes=1)
    assert np.abs(ev / ev_o - 1).max() < 1e-4
    assert pca.subspace_angle(sc, sc_o) < 1e-3


def test_eigensnp_item_passes_with_128_byte_boxes(gpu_ctx, monkeypatch):
    """The batched per-LD-block passes with 128-byte TMA boxes (the shape the condensed-feature pass takes on matrices
    with a large row pitch) against 64-byte boxes: exact integer accumulation on the same fields -> bit-identical
    results; block sizes from 1 SNP to several stages, none aligned to a box."""
    import genomic_pca_b200 as gp
    S = _prep(gpu_ctx, 1000, 5000, 5, seed=52)
    d = S.shape[0]
    sizes = [65, 63, 1, 255, 257, 300, 130, 5, 700, 412, 412, 412, 412, 100, 1025]
    edges = [0]
    for sz in sizes:
        if edges[-1] + sz < d:
            edges.append(edges[-1] + sz)
    edges.append(d)
    blocks = [np.arange(edges[i], edges[i + 1]) for i in range(len(edges) - 1)]
    cfg = gp.EigenSnpConfig(target_num_global_pcs=4, components_per_ld_block=7, subset_factor=0.5, min_subset_size=300,
                            max_subset_size=700, local_oversampling=6, global_oversampling=8, random_seed=21,
                            refine_pass_count=1)
    gpu_ctx.set_sketch_engine(2)
    gpu_ctx.set_batch_blocks(True)
    monkeypatch.setenv("GPCA_I8_ITEM_BOX", "64")
    r64 = gpu_ctx.eigensnp(blocks, cfg)
    monkeypatch.setenv("GPCA_I8_ITEM_BOX", "128")
    r128 = gpu_ctx.eigensnp(blocks, cfg)
    assert all(np.array_equal(a, b) for a, b in zip(r64, r128))
    monkeypatch.setenv("GPCA_DEBUG_NO_ID_ORDER", "1")          # ... and on the slot-ordered copies
    r128s = gpu_ctx.eigensnp(blocks, cfg)
    monkeypatch.setenv("GPCA_I8_ITEM_BOX", "64")
    r64s = gpu_ctx.eigensnp(blocks, cfg)
    assert all(np.array_equal(a, b) for a, b in zip(r64s, r128s))
