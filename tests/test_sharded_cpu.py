"""CPU test (gloo, world_size 2) of the N>1 formulation of the path: SNP-sharded rfit where the only
exchanges are (a) the sum of the N x l sample-side sketch and (b) the sum of l x l Gram matrices.
The numpy mirror below follows csrc/drivers.cu::gpca_rfit step by step (same Philox Omega rows by global
variant index, eigen-based orthonormalisation, no snp-side orthonormalisation between half-steps); run on
two gloo ranks it must reproduce the unsharded oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pca, rng, synth, bed


def _orth_eig(y, allreduce=None):
    for eps in (1e-11, 1e-13):
        g = y.T @ y
        if allreduce is not None:
            g = allreduce(g)
        w, v = np.linalg.eigh(g)
        w, v = w[::-1], v[:, ::-1]
        t = np.where(w > eps * w[0], 1.0 / np.sqrt(np.where(w > 0, w, 1.0)), 0.0)
        y = y @ (v * t)
    return y


def rfit_sharded(S_loc, k, oversample, seed, power_iters, row0, d_total, allreduce, orth_snp_side=False):
    """Mirror of gpca_rfit for one shard S_loc [D_loc, N]; allreduce(x) returns the sum over shards.
    orth_snp_side: the variant gpca_rfit takes when a device holds fewer SNPs than samples -- the D-side iterate is
    re-orthonormalised (its l x l Gram summed over the shards) and the N-side one is left as it comes."""
    d_loc, n = S_loc.shape
    l = min(k + oversample, n, d_total)
    omega = rng.gaussian_matrix(seed, pca.STREAM_RFIT_OMEGA, row0, d_loc, l)
    y = allreduce(S_loc.T @ omega)
    for _ in range(power_iters):
        if orth_snp_side:
            z = _orth_eig(S_loc @ y, allreduce)
        else:
            z = S_loc @ _orth_eig(y)
        y = allreduce(S_loc.T @ z)
    q = _orth_eig(y)
    b = S_loc @ q
    w, vb = np.linalg.eigh(allreduce(b.T @ b))
    w, vb = w[::-1], vb[:, ::-1]
    rot = (b @ vb[:, :k]) / np.sqrt(w[:k])
    scores = allreduce(S_loc.T @ rot)
    scores, rot = pca.fix_signs(scores, rot)
    return scores, w[:k] / (n - 1), rot


def _make(n=300, m=2400):
    g, _ = synth.balding_nichols(n, m, n_pops=5, seed=3)
    keep, mean, sd, _ = bed.snp_qc_and_std_params(g, max_hwe_p=1.0)
    return pca.standardize_dense(g[keep], mean[keep].astype(np.float64), sd[keep].astype(np.float64))


def _worker(rank, world, port, out, shape=(300, 2400), orth_snp_side=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = _make(*shape)
    d = S.shape[0]
    bounds = [d * r // world for r in range(world + 1)]
    lo, hi = bounds[rank], bounds[rank + 1]

    def allreduce(x):
        t = torch.from_numpy(np.ascontiguousarray(x))
        dist.all_reduce(t)
        return t.numpy()

    sc, ev, rot = rfit_sharded(S[lo:hi], 4, 10, 42, 2, lo, d, allreduce, orth_snp_side)
    np.savez(os.path.join(out, f"r{rank}.npz"), sc=sc, ev=ev, rot=rot, lo=lo, hi=hi)
    dist.destroy_process_group()


def test_snp_sharded_rfit_two_ranks_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    S = _make()
    r0 = np.load(tmp_path / "r0.npz")
    r1 = np.load(tmp_path / "r1.npz")
    # replicated outputs identical on both ranks; loadings rows stay local to the shard
    assert np.allclose(r0["sc"], r1["sc"], rtol=0, atol=1e-9)
    assert np.array_equal(r0["ev"], r1["ev"])
    rot = np.concatenate([r0["rot"], r1["rot"]])
    assert rot.shape[0] == S.shape[0]
    # same answer as one shard (same Omega rows by global index), and as the oracle with Householder QR
    one = rfit_sharded(S, 4, 10, 42, 2, 0, S.shape[0], lambda x: x)
    assert np.abs(r0["ev"] / one[1] - 1).max() < 1e-10
    assert pca.subspace_angle(r0["sc"], one[0]) < 1e-7
    sc_o, ev_o, rot_o = pca.rfit(S, 4, 10, seed=42, power_iters=2)
    assert np.abs(r0["ev"] / ev_o - 1).max() < 1e-8
    assert pca.subspace_angle(r0["sc"], sc_o) < 1e-6
    assert pca.subspace_angle(rot, rot_o) < 1e-6


def test_snp_sharded_rfit_snp_side_orthonormalisation_gloo(tmp_path):
    """More samples than SNPs per device: the sharded D-side iterate is orthonormalised through an l x l Gram
    allreduce (drivers.cu, orth_snp_side).  Two gloo ranks against one shard and against the oracle."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    shape = (900, 700)
    mp.spawn(_worker, args=(2, port, str(tmp_path), shape, True), nprocs=2, join=True)
    S = _make(*shape)
    assert S.shape[0] < S.shape[1]
    r0 = np.load(tmp_path / "r0.npz")
    r1 = np.load(tmp_path / "r1.npz")
    assert np.allclose(r0["sc"], r1["sc"], rtol=0, atol=1e-9)
    assert np.array_equal(r0["ev"], r1["ev"])
    rot = np.concatenate([r0["rot"], r1["rot"]])
    one = rfit_sharded(S, 4, 10, 42, 2, 0, S.shape[0], lambda x: x, True)
    assert np.abs(r0["ev"] / one[1] - 1).max() < 1e-10
    assert pca.subspace_angle(r0["sc"], one[0]) < 1e-7
    # the oracle orthonormalises the sample side: same subspace
    sc_o, ev_o, rot_o = pca.rfit(S, 4, 10, seed=42, power_iters=2)
    assert np.abs(r0["ev"] / ev_o - 1).max() < 1e-8
    assert pca.subspace_angle(r0["sc"], sc_o) < 1e-6
    assert pca.subspace_angle(rot, rot_o) < 1e-6
