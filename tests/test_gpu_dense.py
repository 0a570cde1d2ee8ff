"""The split-bf16 tcgen05 product with a dense fp32 matrix (EigenSNP's global randomized SVD on the condensed
features; replaces the cuBLAS SGEMMs of round 1) against a float64 product of the same operands."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cols_mode", [False, True])
@pytest.mark.parametrize("shape", [(5000, 1484, 30), (777, 260, 7), (40000, 131, 32), (300, 9000, 20)])
def test_dense_product_matches_f64(gpu_ctx, cols_mode, shape):
    import torch
    n, r, l = shape
    ldc = -(-r // 4) * 4 + 4                       # padded row stride (multiple of 4 floats)
    g = torch.Generator(device="cuda").manual_seed(n + r + l)
    cmat = torch.zeros((n, ldc), device="cuda")
    # condensed features are O(1) with a common offset: the offset exercises the rank-one (mean) term
    cmat[:, :r] = torch.randn((n, r), device="cuda", generator=g) * 3.0 + 1.5
    k, rows = (n, r) if cols_mode else (r, n)
    ld = l + 3
    w = torch.zeros((k, ld), device="cuda")
    w[:, :l] = torch.randn((k, l), device="cuda", generator=g)
    f = torch.rand(k, device="cuda", generator=g) + 0.5
    e = torch.randn(k, device="cuda", generator=g)
    a = torch.rand(rows, device="cuda", generator=g) + 0.5
    b = torch.randn(rows, device="cuda", generator=g)
    out = torch.full((rows, l + 1), 7.0, device="cuda")
    torch.cuda.synchronize()
    gpu_ctx.dense_product(cmat.data_ptr(), n, r, ldc, cols_mode, w.data_ptr(), l, ld, out.data_ptr(), l + 1,
                          f=f.data_ptr(), e=e.data_ptr(), a=a.data_ptr(), b=b.data_ptr())
    gpu_ctx.synchronize()
    x = cmat[:, :r].double()
    x = x.t() if cols_mode else x
    wd = w[:, :l].double()
    ref = a.double()[:, None] * (x @ (f.double()[:, None] * wd)) - b.double()[:, None] * (e.double() @ wd)[None, :]
    got = out[:, :l].double()
    scale = (x.abs() @ (f.double()[:, None] * wd).abs()).max()          # size of the sums being formed
    assert float((got - ref).abs().max() / scale) < 3e-5
    assert torch.all(out[:, l] == 7.0)                                  # columns past l are not touched
    # without the optional vectors
    gpu_ctx.dense_product(cmat.data_ptr(), n, r, ldc, cols_mode, w.data_ptr(), l, ld, out.data_ptr(), l + 1)
    gpu_ctx.synchronize()
    ref = x @ wd - wd.sum(0)[None, :]
    assert float((out[:, :l].double() - ref).abs().max() / scale) < 3e-5
