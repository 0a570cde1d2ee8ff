"""Host-side logic of bench.py that does not need a GPU: the strong-scaling split of BASELINE config 4 over the ranks
(whole LD blocks, every SNP exactly once) and the number formatting helpers it shares with the tests."""
import numpy as np

import bench


def test_strong_scaling_shards_cover_every_snp_once():
    m, nb = bench.C4_SNPS, bench.C4_BLOCKS
    edges = np.linspace(0, m, nb + 1).astype(np.int64)
    for world in (1, 2, 3, 4, 8):
        seen_blocks, next_snp = 0, 0
        for rank in range(world):
            b0, b1, s0, s1 = bench.shard_of(rank, world, nb, m)
            assert b0 == seen_blocks and s0 == next_snp            # contiguous, in rank order
            assert s0 == edges[b0] and s1 == edges[b1]             # cut at LD-block boundaries only
            assert b1 > b0 and s1 > s0
            seen_blocks, next_snp = b1, s1
        assert seen_blocks == nb and next_snp == m
        sizes = [bench.shard_of(r, world, nb, m)[3] - bench.shard_of(r, world, nb, m)[2] for r in range(world)]
        assert max(sizes) - min(sizes) <= 2 * (m // nb + 1)        # balanced to within a block or two


def test_subspace_angle_helper():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((200, 5))
    q, _ = np.linalg.qr(a)
    assert bench.subspace_angle(a, q @ rng.standard_normal((5, 5))) < 1e-7          # same span
    b = q.copy()
    b[:, 0] += 1e-3 * rng.standard_normal(200)
    ang = bench.subspace_angle(q, b)
    assert 1e-4 < ang < 5e-2


def test_no_collective_call_under_a_rank_condition():
    """bench.py under torchrun: every rank must issue the same collectives.  gpca_rfit, gpca_eigensnp and the sample-side
    sketch of a sharded context end in an exchange, and the timing helpers reduce over the ranks -- none of them may sit
    under a condition on the rank (a check that rank 0 ran alone once left the other ranks waiting for ever).  The
    full-size parity check, which calls the sample-side sketch, must be tied to one-GPU runs."""
    import ast
    import inspect
    collective = {"rfit", "eigensnp", "sketch_sample_side", "comm_init", "sampled_parity", "timed_rfit", "e2e_loop",
                  "max_over_ranks", "sum_over_ranks", "barrier", "broadcast_object_list", "all_reduce"}
    tree = ast.parse(inspect.getsource(bench.run_ours))

    def names(node):
        return {n.id for n in ast.walk(node) if isinstance(n, ast.Name)}

    def calls(nodes):
        out = set()
        for node in nodes:
            for c in ast.walk(node):
                if isinstance(c, ast.Call):
                    f = c.func
                    out.add(f.attr if isinstance(f, ast.Attribute) else getattr(f, "id", ""))
        return out

    checked = parity_guard = 0
    for node in ast.walk(tree):
        if isinstance(node, ast.If):
            if "rank" in names(node.test):
                checked += 1
                bad = calls(node.body) & collective
                # (leaving the process group on the way out is the one rank-dependent collective-free exit)
                assert not bad, f"collective call(s) {bad} under `if {ast.unparse(node.test)}`"
            if "sampled_parity" in calls(node.body):
                parity_guard += 1
                assert ast.unparse(node.test).startswith("world == 1"), ast.unparse(node.test)
    assert checked >= 2 and parity_guard == 1


def test_watchdog_ends_a_run_that_never_finishes():
    import subprocess
    import sys
    r = subprocess.run([sys.executable, "-c", "import time, bench; bench.start_watchdog(0.5); time.sleep(30)"],
                       capture_output=True, text=True, cwd=bench.os.path.dirname(bench.os.path.abspath(bench.__file__)))
    assert r.returncode == 3 and "giving up" in r.stderr
    r = subprocess.run([sys.executable, "-c", "import bench; assert bench.start_watchdog(0) is None"],
                       capture_output=True, text=True, cwd=bench.os.path.dirname(bench.os.path.abspath(bench.__file__)))
    assert r.returncode == 0, r.stderr
