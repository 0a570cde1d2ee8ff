"""GPU parity of the sketch passes against a dense f64 product with the standardized matrix."""
import numpy as np
import pytest
import torch

from helpers import make_dataset, standardized

pytestmark = pytest.mark.gpu


def _setup(ctx, n, m, seed, missing_rate=0.0, call_rate=0.9):
    import genomic_pca_b200 as gp
    g, payload = make_dataset(n, m, seed=seed, missing_rate=missing_rate)
    ctx.load_bed(payload, n, m)
    keep, mean, sd, _ = ctx.snp_qc(gp.QcConfig(call_rate, 0.01, 1.0))
    idx = np.nonzero(keep)[0]
    ctx.set_pca_snps(idx, mean[idx], sd[idx])
    return standardized(g[idx], mean[idx], sd[idx])


def _run(ctx, S, l, engine, seed=0):
    d, n = S.shape
    r = np.random.default_rng(seed)
    ctx.set_sketch_engine(engine)
    y = r.standard_normal((n, l)).astype(np.float32)
    w = r.standard_normal((d, l)).astype(np.float32)
    ty = torch.from_numpy(y).cuda()
    tw = torch.from_numpy(w).cuda()
    out_d = torch.full((d, l), float("nan"), device="cuda", dtype=torch.float32)
    out_n = torch.full((n, l), float("nan"), device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    ctx.sketch_snp_side(ty.data_ptr(), out_d.data_ptr(), l, l)
    ctx.sketch_sample_side(tw.data_ptr(), out_n.data_ptr(), l, l)
    ctx.synchronize()
    ref_d = S @ y.astype(np.float64)
    ref_n = S.T @ w.astype(np.float64)
    return out_d.cpu().numpy(), ref_d, out_n.cpu().numpy(), ref_n


def _relerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("n,m,l", [(64, 200, 5), (130, 700, 30), (2504, 3000, 30), (1000, 257, 64), (33, 40, 1),
                                   (5000, 600, 17)])
def test_sketch_simt_matches_dense(gpu_ctx, n, m, l):
    S = _setup(gpu_ctx, n, m, seed=n + m)
    od, rd, on, rn = _run(gpu_ctx, S, l, engine=0)
    assert _relerr(od, rd) < 2e-5          # fp32 accumulation vs f64
    assert _relerr(on, rn) < 2e-5


def test_sketch_with_missing_calls_mean_imputed(gpu_ctx):
    S = _setup(gpu_ctx, 300, 500, seed=3, missing_rate=0.02)
    od, rd, on, rn = _run(gpu_ctx, S, 12, engine=0)
    assert _relerr(od, rd) < 2e-5
    assert _relerr(on, rn) < 2e-5


def test_sketch_linearity_full_width(gpu_ctx):
    """Size-independent property: S(aY1 + Y2) = a S Y1 + S Y2 (checked at a larger shape than the oracle needs)."""
    S = _setup(gpu_ctx, 20000, 4000, seed=8)
    d, n = S.shape
    l = 30
    gpu_ctx.set_sketch_engine(0)
    g = torch.Generator(device="cuda").manual_seed(1)
    y1 = torch.randn(n, l, device="cuda", generator=g)
    y2 = torch.randn(n, l, device="cuda", generator=g)
    o1 = torch.empty(d, l, device="cuda")
    o2 = torch.empty(d, l, device="cuda")
    o3 = torch.empty(d, l, device="cuda")
    y3 = (0.5 * y1 + y2).contiguous()
    torch.cuda.synchronize()
    for a, b in ((y1, o1), (y2, o2), (y3, o3)):
        gpu_ctx.sketch_snp_side(a.data_ptr(), b.data_ptr(), l, l)
    gpu_ctx.synchronize()
    err = (o3 - (0.5 * o1 + o2)).abs().max() / o3.abs().max()
    assert err < 1e-4
    # adjointness: <S y, w> == <y, S^T w>
    w = torch.randn(d, l, device="cuda", generator=g)
    ow = torch.empty(n, l, device="cuda")
    torch.cuda.synchronize()
    gpu_ctx.sketch_sample_side(w.data_ptr(), ow.data_ptr(), l, l)
    gpu_ctx.synchronize()
    lhs = (o1.double() * w.double()).sum()
    rhs = (y1.double() * ow.double()).sum()
    assert abs(lhs - rhs) / abs(lhs) < 1e-4


@pytest.mark.parametrize("n,m,l", [(2504, 3000, 30), (700, 1500, 5), (4100, 900, 64), (256, 128 * 3 + 5, 17),
                                   (20000, 2000, 30)])
def test_sketch_tcgen05_matches_dense(gpu_ctx, n, m, l):
    """tcgen05 engine (fp16 operands, fp32 TMEM accumulators): the only rounding is B' -> fp16 (2^-11)."""
    S = _setup(gpu_ctx, n, m, seed=n + m + 1)
    od, rd, on, rn = _run(gpu_ctx, S, l, engine=1)
    assert _relerr(od, rd) < 1.5e-3
    assert _relerr(on, rn) < 1.5e-3
    # and it agrees with the SIMT engine to the same level
    od0, _, on0, _ = _run(gpu_ctx, S, l, engine=0)
    assert _relerr(od, od0) < 1.5e-3 and _relerr(on, on0) < 1.5e-3


def test_sketch_tcgen05_with_missing(gpu_ctx):
    S = _setup(gpu_ctx, 1500, 1200, seed=5, missing_rate=0.02)
    od, rd, on, rn = _run(gpu_ctx, S, 20, engine=1)
    assert _relerr(od, rd) < 1.5e-3
    assert _relerr(on, rn) < 1.5e-3


@pytest.mark.parametrize("n,m,l", [(2504, 3000, 30), (700, 1500, 5), (256, 128 * 3 + 5, 17), (20000, 2000, 30),
                                   (1000, 40000, 32)])
def test_sketch_int8_engine_matches_dense(gpu_ctx, n, m, l):
    """tcgen05 kind::i8 engine: exact int32 accumulation; the only rounding is the 16-bit quantisation of the dense
    operand (relative to its global max)."""
    S = _setup(gpu_ctx, n, m, seed=n + m + 2)
    od, rd, on, rn = _run(gpu_ctx, S, l, engine=2)
    assert _relerr(od, rd) < 3e-4
    assert _relerr(on, rn) < 3e-4


def test_sketch_int8_engine_missing_and_determinism(gpu_ctx):
    S = _setup(gpu_ctx, 1500, 1200, seed=6, missing_rate=0.02)
    od, rd, on, rn = _run(gpu_ctx, S, 20, engine=2)
    assert _relerr(od, rd) < 3e-4 and _relerr(on, rn) < 3e-4
    od2, _, on2, _ = _run(gpu_ctx, S, 20, engine=2)
    assert np.array_equal(od, od2) and np.array_equal(on, on2)      # integer accumulation: bit-reproducible


def _device_dataset(ctx, n, m):
    """Synthetic .bed payload generated on the device (bench.py's generator): shapes too large for the numpy oracle."""
    import bench
    import genomic_pca_b200 as gp
    dev = torch.device("cuda", 0)
    payload = bench.synth_bed_device(torch, n, m, 0, dev)
    ctx.load_bed_device(payload.data_ptr(), n, m)
    keep, mean, sd, _ = ctx.snp_qc(gp.QcConfig(0.98, 0.0, 1.0))
    return ctx.set_pca_snps_mask(keep, mean, sd)


@pytest.mark.parametrize("engine", [1, 2])
def test_tensor_engines_with_two_ctas_per_sm(gpu_ctx, engine):
    """Shapes with more work items than SMs: two persistent CTAs share every SM and each walks several items.
    (The small parity cases above never co-locate two CTAs.)  Checked against the SIMT engine, and run to run:
    every reduction has a fixed order, so repeated passes must agree bit for bit."""
    n, m, l = 4096, 100_000, 30
    d = _device_dataset(gpu_ctx, n, m)
    dev = torch.device("cuda", 0)
    ext = torch.cuda.ExternalStream(gpu_ctx.stream)
    with torch.cuda.stream(ext):
        g = torch.Generator(device=dev)
        g.manual_seed(5)
        Bs = torch.randn(n, l, device=dev, generator=g)
        Bd = torch.randn(d, l, device=dev, generator=g)
        ext.synchronize()
        for fn, src, rows in ((gpu_ctx.sketch_snp_side, Bs, d), (gpu_ctx.sketch_sample_side, Bd, n)):
            outs = []
            for eng in (0, engine, engine, engine):
                gpu_ctx.set_sketch_engine(eng)
                o = torch.empty(rows, l, device=dev)
                fn(src.data_ptr(), o.data_ptr(), l, l)
                gpu_ctx.synchronize()
                outs.append(o.cpu().numpy())
            ref = outs[0]
            tol = 2e-3 if engine == 1 else 5e-4
            assert np.abs(outs[1] - ref).max() / np.abs(ref).max() < tol
            assert np.array_equal(outs[1], outs[2]) and np.array_equal(outs[1], outs[3])
    gpu_ctx.set_sketch_engine(2)


def test_rfit_reproducible_at_multi_item_scale(gpu_ctx):
    """Same seed, same context -> identical eigenvalues and scores (2 CTAs per SM, several items per CTA)."""
    _device_dataset(gpu_ctx, 4096, 100_000)
    a = gpu_ctx.rfit(10, 10, 2, seed=42)
    b = gpu_ctx.rfit(10, 10, 2, seed=42)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
    gpu_ctx.set_sketch_engine(0)
    c = gpu_ctx.rfit(10, 10, 2, seed=42)
    gpu_ctx.set_sketch_engine(2)
    from oracle import pca
    assert np.abs(a[1] / c[1] - 1).max() < 1e-4
    assert pca.subspace_angle(a[0], c[0]) < 1e-3


@pytest.mark.parametrize("switch", ["GPCA_I8_WIDE", "GPCA_I8_TILE_SYNC", "GPCA_I8_DEEP", "GPCA_I8_WIDE+GPCA_I8_TILE_SYNC",
                                    "GPCA_I8_DEEP+GPCA_I8_TILE_SYNC"])
@pytest.mark.parametrize("ksplit", ["1", "3"])
def test_int8_engine_variants_equal_regular(gpu_ctx, monkeypatch, ksplit, switch):
    """Variants of the integer engine's schedule -- the 512-row CTA shape (one CTA per SM, the whole TMEM: every
    operand-image stage shared by twice as many rows), the per-row-tile TMEM hand-over, both together, and the deep-slot
    shape (one 256-row CTA per SM with six TMEM slots and two expander warps per lane quarter) -- against the regular kernel on
    the same K split: the integer accumulation is exact and the fp32 epilogue is per row, so they must agree bit for bit
    -- odd row counts, several items per CTA, both pass orientations, with and without split-K partials."""
    n, m, l = 4000 + 77, 150_000, 30
    d = _device_dataset(gpu_ctx, n, m)
    dev = torch.device("cuda", 0)
    gpu_ctx.set_sketch_engine(2)
    monkeypatch.setenv("GPCA_DEBUG_KSPLIT", ksplit)
    ext = torch.cuda.ExternalStream(gpu_ctx.stream)
    with torch.cuda.stream(ext):
        g = torch.Generator(device=dev)
        g.manual_seed(11)
        Bs = torch.randn(n, l, device=dev, generator=g)
        Bd = torch.randn(d, l, device=dev, generator=g)
        ext.synchronize()
        for fn, src, rows in ((gpu_ctx.sketch_snp_side, Bs, d), (gpu_ctx.sketch_sample_side, Bd, n)):
            outs = {}
            for wide in ("0", "1", "1"):
                for name in switch.split("+"):
                    monkeypatch.setenv(name, wide)
                o = torch.full((rows, l), float("nan"), device=dev)
                fn(src.data_ptr(), o.data_ptr(), l, l)
                gpu_ctx.synchronize()
                outs.setdefault(wide, []).append(o.cpu().numpy())
            assert np.isfinite(outs["1"][0]).all()
            assert np.array_equal(outs["0"][0], outs["1"][0]) and np.array_equal(outs["1"][0], outs["1"][1])
