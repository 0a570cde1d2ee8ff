"""GPU parity of the rfit driver: against the oracle's restatement with the same Philox test matrix,
and against the exact f64 eigen-decomposition.  Tolerances are the north star's:
eigenvalues 1e-4 relative, principal subspace angles < 1e-3 rad."""
import numpy as np
import pytest

from oracle import pca

from helpers import make_dataset, standardized

pytestmark = pytest.mark.gpu

EV_RTOL = 1e-4
ANGLE_TOL = 1e-3


def _prep(ctx, n, m, n_pops, seed, vcf=False):
    import genomic_pca_b200 as gp
    g, payload = make_dataset(n, m, n_pops=n_pops, seed=seed)
    if vcf:
        ctx.load_u8_variant_major(g.astype(np.uint8))
        keep, mean, sd = ctx.vcf_maf_filter(0.01)
    else:
        ctx.load_bed(payload, n, m)
        keep, mean, sd, _ = ctx.snp_qc(gp.QcConfig(0.98, 0.01, 1.0))
    idx = np.nonzero(keep)[0]
    ctx.set_pca_snps(idx, mean[idx], sd[idx])
    return standardized(g[idx], mean[idx], sd[idx])


@pytest.mark.parametrize("engine", [0, 1, 2])
@pytest.mark.parametrize("n,m,pops,k", [(600, 4000, 5, 4), (2504, 6000, 6, 5)])
def test_rfit_matches_oracle_and_exact(gpu_ctx, engine, n, m, pops, k):
    S = _prep(gpu_ctx, n, m, pops, seed=n, vcf=True)
    gpu_ctx.set_sketch_engine(engine)
    sc, ev, ld = gpu_ctx.rfit(k, 10, power_iters=3, seed=42)
    sc_o, ev_o, ld_o = pca.rfit(S, k, 10, seed=42, power_iters=3)
    sc_x, ev_x, ld_x = pca.exact_pca(S, k)
    assert np.abs(ev / ev_o - 1).max() < EV_RTOL
    assert np.abs(ev / ev_x - 1).max() < EV_RTOL
    assert pca.subspace_angle(sc, sc_o) < ANGLE_TOL
    assert pca.subspace_angle(ld, ld_o) < ANGLE_TOL
    assert pca.subspace_angle(sc, sc_x) < 3e-3     # limited by the randomized method itself (oracle shows the same)
    # sign convention + column-wise agreement with the oracle
    assert np.abs(sc - sc_o).max() / np.abs(sc_o).max() < 5e-3
    # scores-only call: the rotation is never formed (scores = (S^T B) T), same result
    sc2, ev2, none = gpu_ctx.rfit(k, 10, power_iters=3, seed=42, want_loadings=False)
    assert none is None and np.array_equal(ev2, ev)
    assert pca.subspace_angle(sc2, sc_o) < ANGLE_TOL
    assert np.abs(sc2 - sc).max() / np.abs(sc).max() < 2e-3


def test_rfit_reference_argument_rules(gpu_ctx):
    import genomic_pca_b200 as gp
    S = _prep(gpu_ctx, 40, 300, 3, seed=1)
    with pytest.raises(gp.GpcaError):
        gpu_ctx.rfit(0)                                   # main.rs:607
    sc, ev, ld = gpu_ctx.rfit(60, 10, seed=1)             # k capped at min(N, D) = 40 (main.rs:621-628)
    assert sc.shape[1] == 40 and ev.shape[0] == 40
    # seeded runs are reproducible, unseeded runs still valid
    a = gpu_ctx.rfit(3, 10, seed=7)[1]
    b = gpu_ctx.rfit(3, 10, seed=7)[1]
    assert np.array_equal(a, b)
    c = gpu_ctx.rfit(3, 10, seed=None)[1]
    # 3 populations -> 2 structured components; the third sits in the noise bulk and depends on Omega
    assert np.all(np.isfinite(c)) and np.abs(c[:2] / a[:2] - 1).max() < 0.05 and abs(c[2] / a[2] - 1) < 0.3


def test_results_download_in_chunks(gpu_ctx, monkeypatch):
    """Scores (widened to f64 on host threads) and loadings leave the device through pinned landing buffers in chunks;
    a forced chunk of 1,000 floats (dozens of chunks, both buffers in use, ragged tail) must give the same arrays as the
    single-chunk path, bit for bit."""
    _prep(gpu_ctx, 700, 3000, 4, seed=3)
    monkeypatch.delenv("GPCA_DEBUG_DOWNLOAD_CHUNK", raising=False)
    sc1, ev1, ld1 = gpu_ctx.rfit(6, 10, power_iters=2, seed=11, want_loadings=True)
    monkeypatch.setenv("GPCA_DEBUG_DOWNLOAD_CHUNK", "1000")
    sc2, ev2, ld2 = gpu_ctx.rfit(6, 10, power_iters=2, seed=11, want_loadings=True)
    assert sc1.dtype == np.float64 and ld1.dtype == np.float32
    assert np.array_equal(sc1, sc2) and np.array_equal(ld1, ld2) and np.array_equal(ev1, ev2)
    # the f64 scores are exactly the widened fp32 values
    assert np.array_equal(sc1, sc1.astype(np.float32).astype(np.float64))


@pytest.mark.parametrize("k,oversample", [(24, 16), (40, 10), (12, 22)])
def test_rfit_wide_sketch_matches_oracle(gpu_ctx, k, oversample):
    """l = k + oversample in (32, 64]: the shape of BASELINE config 5 (k = 40, l = 50).  The default engine hands these
    to the fp16 tensor engine (64 columns), and the N-side helpers run their 64-column variants (Gram tiles split over
    3 to 8 warps per row group, Y.T with 12 / 16 owned columns)."""
    S = _prep(gpu_ctx, 1500, 5000, k + 2, seed=100 + k, vcf=True)
    gpu_ctx.set_sketch_engine(2)
    sc, ev, ld = gpu_ctx.rfit(k, oversample, power_iters=2, seed=5)
    sc_o, ev_o, ld_o = pca.rfit(S, k, oversample, seed=5, power_iters=2)
    assert sc.shape == (1500, k) and ld.shape == (S.shape[0], k)
    assert np.abs(ev / ev_o - 1).max() < EV_RTOL
    assert pca.subspace_angle(sc, sc_o) < ANGLE_TOL
    assert pca.subspace_angle(ld, ld_o) < ANGLE_TOL
    # run to run
    sc2, ev2, ld2 = gpu_ctx.rfit(k, oversample, power_iters=2, seed=5)
    assert np.array_equal(ev, ev2) and np.array_equal(sc, sc2) and np.array_equal(ld, ld2)


def test_rfit_generated_test_matrix_matches_materialised(gpu_ctx, monkeypatch):
    """The first pass quantises the Gaussian test matrix straight from the Philox generator (a-priori scale from the
    generator's |z| bound and max 1/sd) instead of writing it to memory, measuring its max and reading it back.  Same
    normals, slightly coarser quantisation step: the two paths agree far inside the parity tolerances."""
    S = _prep(gpu_ctx, 1300, 9000, 6, seed=77, vcf=True)
    gpu_ctx.set_sketch_engine(2)
    monkeypatch.delenv("GPCA_DEBUG_NO_GEN_FUSE", raising=False)
    sc, ev, ld = gpu_ctx.rfit(5, 10, power_iters=2, seed=9)
    monkeypatch.setenv("GPCA_DEBUG_NO_GEN_FUSE", "1")
    sc_m, ev_m, ld_m = gpu_ctx.rfit(5, 10, power_iters=2, seed=9)
    assert np.abs(ev / ev_m - 1).max() < 2e-5      # (the north-star tolerance is 1e-4)
    assert pca.subspace_angle(sc, sc_m) < 2e-4 and pca.subspace_angle(ld, ld_m) < 2e-4
    # with no power iteration the estimate depends on every rounding of the test matrix (a q = 0 randomized SVD is only
    # accurate to ~1e-3 on the trailing components here): the two quantisations stay inside that
    monkeypatch.delenv("GPCA_DEBUG_NO_GEN_FUSE", raising=False)
    sc0, ev0, _ = gpu_ctx.rfit(5, 10, power_iters=0, seed=9)
    monkeypatch.setenv("GPCA_DEBUG_NO_GEN_FUSE", "1")
    sc0_m, ev0_m, _ = gpu_ctx.rfit(5, 10, power_iters=0, seed=9)
    assert np.abs(ev0 / ev0_m - 1).max() < 2e-3
    assert pca.subspace_angle(sc0[:, :3], sc0_m[:, :3]) < 5e-3
    sc_o, ev_o, _ = pca.rfit(S, 5, 10, seed=9, power_iters=0)
    assert np.abs(ev0 / ev_o - 1).max() < 2e-3 and np.abs(ev0_m / ev_o - 1).max() < 2e-3


@pytest.mark.parametrize("engine", [0, 2])
def test_rfit_more_samples_than_snps(gpu_ctx, engine, monkeypatch):
    """N > D: the power iteration re-orthonormalises the SNP side (the one with fewer rows) instead of the sample side.
    Same subspace either way: parity with the oracle (which orthonormalises the sample side) and with the other
    ordering (GPCA_DEBUG_ORTH_SAMPLE_SIDE)."""
    S = _prep(gpu_ctx, 3000, 1400, 5, seed=17, vcf=True)
    assert S.shape[0] < 3000
    gpu_ctx.set_sketch_engine(engine)
    monkeypatch.delenv("GPCA_DEBUG_ORTH_SAMPLE_SIDE", raising=False)
    sc, ev, ld = gpu_ctx.rfit(4, 10, power_iters=2, seed=3)
    sc_o, ev_o, ld_o = pca.rfit(S, 4, 10, seed=3, power_iters=2)
    assert np.abs(ev / ev_o - 1).max() < EV_RTOL
    assert pca.subspace_angle(sc, sc_o) < ANGLE_TOL and pca.subspace_angle(ld, ld_o) < ANGLE_TOL
    monkeypatch.setenv("GPCA_DEBUG_ORTH_SAMPLE_SIDE", "1")
    sc_s, ev_s, ld_s = gpu_ctx.rfit(4, 10, power_iters=2, seed=3)
    assert np.abs(ev / ev_s - 1).max() < EV_RTOL
    assert pca.subspace_angle(sc, sc_s) < ANGLE_TOL and pca.subspace_angle(ld, ld_s) < ANGLE_TOL
