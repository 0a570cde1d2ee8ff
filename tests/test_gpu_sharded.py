"""The SNP-sharded drivers on ONE GPU: two contexts in two host threads, each holding half of the SNPs / LD blocks, with
a gpca_allreduce_fn that adds the two device buffers behind a host barrier.  This drives the real sharded code paths
of the library (gpca_set_shard, the exchange after every sample-side pass, the l x l Gram exchange of a sharded
orthonormalisation, the rank-invariant choice of the re-orthonormalised side) and compares with the unsharded run --
what tools/multi_gpu_check.py does with NCCL on two GPUs."""
import threading

import numpy as np
import pytest

from oracle import pca

from helpers import make_dataset

pytestmark = pytest.mark.gpu


class TwoShardExchange:
    """Sum of the two shards' buffers.  Each shard's hook first drains its own stream (its partial is complete), both
    meet at a barrier, shard 0 adds the buffers with torch on the default stream and writes the sum to both, and a
    second barrier releases them.  Host-side waiting only: no kernel ever waits for another launch."""

    def __init__(self):
        self.barrier = threading.Barrier(2)
        self.args = [None, None]
        self.calls = 0

    def hook(self, rank, ctx):
        import torch

        def view(ptr, count, dtype):
            iface = {"shape": (count,), "typestr": "<f4" if dtype == 0 else "<f8", "data": (ptr, False), "version": 2}
            holder = type("P", (), {"__cuda_array_interface__": iface})()
            return torch.as_tensor(holder, device="cuda:0")

        def fn(ptr, count, dtype, stream):
            ctx.synchronize()
            self.args[rank] = (ptr, count, dtype)
            self.barrier.wait(timeout=120)
            if rank == 0:
                assert self.args[0][1:] == self.args[1][1:], "the shards issued different collectives"
                a, b = view(*self.args[0]), view(*self.args[1])
                s = a + b
                a.copy_(s)
                b.copy_(s)
                torch.cuda.synchronize()
                self.calls += 1
            self.barrier.wait(timeout=120)
        return fn


def _run_two_shards(work):
    """work(rank, exchange) -> result, run in two threads; exceptions are re-raised in the caller"""
    ex = TwoShardExchange()
    out, err = [None, None], [None, None]

    def body(r):
        try:
            out[r] = work(r, ex)
        except BaseException as e:      # noqa: BLE001 -- a failing shard must not leave the other at the barrier
            err[r] = e
            ex.barrier.abort()

    th = [threading.Thread(target=body, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in err:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in err:
        if e is not None:
            raise e
    return out, ex


@pytest.mark.parametrize("shape", [(900, 6000), (6000, 2400)])     # D_shard > N (sample side orthonormalised) and < N (SNP side)
def test_sharded_rfit_on_one_gpu_equals_unsharded(shape):
    import genomic_pca_b200 as gp
    n, m = shape
    g, payload = make_dataset(n, m, n_pops=5, seed=1234 + n)
    qc = gp.QcConfig(0.9, 0.01, 1.0)
    full = gp.Context(0)
    keep, mean, sd, _, d = full.ingest_bed(payload, n, m, qc=qc)
    sc0, ev0, ld0 = full.rfit(4, 8, 2, seed=42)
    full.close()
    split = m // 2 + 37                                # unequal shards
    kept_before = [0, int(keep[:split].sum())]
    rows = [(0, split), (split, m)]

    def work(r, ex):
        ctx = gp.Context(0)
        a, b = rows[r]
        _, _, _, _, dr = ctx.ingest_bed(payload[a:b], n, b - a, qc=qc)
        ctx.set_shard(kept_before[r], d)
        ctx.set_allreduce(ex.hook(r, ctx))
        res = ctx.rfit(4, 8, 2, seed=42)
        with pytest.raises(gp.GpcaError):              # every shard must sketch with the same test matrix
            ctx.rfit(4, 8, 2, seed=None)
        ctx.close()
        return res

    (r0, r1), ex = _run_two_shards(work)
    assert ex.calls >= 5                               # the exchange really ran (4 sample-side passes + the side vote)
    for r in (r0, r1):
        assert np.abs(r[1] / ev0 - 1).max() < 1e-4
        assert pca.subspace_angle(r[0], sc0) < 1e-3
    assert np.array_equal(r0[0], r1[0]) and np.array_equal(r0[1], r1[1])     # both shards hold the same scores
    load = np.concatenate([r0[2], r1[2]])
    assert load.shape == ld0.shape and pca.subspace_angle(load, ld0) < 1e-3


def test_sharded_eigensnp_on_one_gpu_equals_unsharded():
    import genomic_pca_b200 as gp
    n, m = 1200, 5200
    g, payload = make_dataset(n, m, n_pops=5, seed=77)
    qc = gp.QcConfig(0.9, 0.0, 1.0)                    # keep every SNP: shard-local ids are offsets of the global ones
    cfg = gp.EigenSnpConfig(target_num_global_pcs=4, components_per_ld_block=5, subset_factor=0.5, min_subset_size=300,
                            max_subset_size=700, local_oversampling=6, global_oversampling=8, random_seed=5,
                            refine_pass_count=1)
    split = 2600
    edges = [list(range(0, split, 325)) + [split], list(range(split, m, 433)) + [m]]
    full = gp.Context(0)
    _, _, _, _, d = full.ingest_bed(payload, n, m, qc=qc)
    assert d == m
    blocks_full = [np.arange(e[i], e[i + 1]) for e in edges for i in range(len(e) - 1)]
    sc0, ev0, ld0 = full.eigensnp(blocks_full, cfg)
    full.close()

    def work(r, ex):
        ctx = gp.Context(0)
        a, b = (0, split) if r == 0 else (split, m)
        ctx.ingest_bed(payload[a:b], n, b - a, qc=qc)
        ctx.set_shard(a, m)
        ctx.set_allreduce(ex.hook(r, ctx))
        e = edges[r]
        res = ctx.eigensnp([np.arange(e[i] - a, e[i + 1] - a) for i in range(len(e) - 1)], cfg)
        ctx.close()
        return res

    (r0, r1), ex = _run_two_shards(work)
    assert ex.calls >= 4
    for r in (r0, r1):
        assert np.abs(r[1] / ev0 - 1).max() < 1e-4
        assert pca.subspace_angle(r[0], sc0) < 1e-3
    load = np.concatenate([r0[2], r1[2]])
    assert pca.subspace_angle(load, ld0) < 1e-3
