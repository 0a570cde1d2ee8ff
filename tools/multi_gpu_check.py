"""Sharded (one process per GPU, the library's own NCCL communicator) rfit and EigenSNP against the same computation on
one GPU.
Run:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_check.py
Every rank loads its contiguous shard of SNPs (whole LD blocks); rank 0 also runs the unsharded problem and compares:
eigenvalues 1e-4 relative, score subspace angle < 1e-3 rad (the north-star tolerances)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench, genomic_pca_b200 as gp
from oracle import pca

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
# default: more SNPs than samples per device; `python ... multi_gpu_check.py 20000 8000 20` covers the other case (the
# rfit power iteration then orthonormalises the sharded SNP side through the Gram allreduce)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
m_shard = int(sys.argv[2]) if len(sys.argv) > 2 else 24_000
nblk_shard = int(sys.argv[3]) if len(sys.argv) > 3 else 40
qc = gp.QcConfig(0.98, 0.0, 1.0)                 # keep every SNP: shards and the full run see the same set


full_payload = bench.synth_bed_device(torch, n, world * m_shard, 0, dev)     # every rank generates the same matrix


def load(ctx, snp0, m):
    payload = full_payload[snp0:snp0 + m].contiguous()                       # ... and takes its rows
    torch.cuda.synchronize()
    ctx.load_bed_device(payload.data_ptr(), n, m)
    keep, mean, sd, _ = ctx.snp_qc(qc)
    return ctx.set_pca_snps_mask(keep, mean, sd)


def blocks_for(d, nblk):
    edges = np.linspace(0, d, nblk + 1).astype(np.int64)
    return [np.arange(edges[i], edges[i + 1], dtype=np.uint64) for i in range(nblk)]


ctx = gp.Context(local)
box = [gp.binding.comm_unique_id() if rank == 0 else None]       # torch.distributed only carries the id bytes
dist.broadcast_object_list(box, src=0)
ctx.comm_init(box[0], rank, world)
assert ctx.comm_world == world
ctx.set_shard(rank * m_shard, world * m_shard)
d = load(ctx, rank * m_shard, m_shard)
assert d == m_shard
sc_r, ev_r, _ = ctx.rfit(8, 10, power_iters=2, seed=42, want_loadings=False)
cfg = gp.EigenSnpConfig(target_num_global_pcs=6, min_subset_size=1500, max_subset_size=3000, subset_factor=0.4)
sc_e, ev_e, ld_e = ctx.eigensnp(blocks_for(d, nblk_shard), cfg)
n_coll = ctx.collective_count
sc_n, ev_n, _ = ctx.rfit(8, 10, power_iters=2, seed=None, want_loadings=False)     # entropy seed: broadcast from rank 0
ctx.close()
# the entropy seed came from rank 0: every rank must have sketched with the same test matrix -> identical scores
same = [None] * world
dist.all_gather_object(same, (ev_n.tobytes(), sc_n.tobytes()))
entropy_seed_ranks_agree = all(x == same[0] for x in same)
dist.barrier()
out = {"world": world, "n": n, "m_shard": m_shard, "collectives": n_coll}
if rank == 0:
    full = gp.Context(local)
    dfull = load(full, 0, world * m_shard)
    sc0, ev0, _ = full.rfit(8, 10, power_iters=2, seed=42, want_loadings=False)
    # the same blocks as the shards used, in global PcaSnpId numbering
    blk = []
    for r in range(world):
        blk += [b + np.uint64(r * m_shard) for b in blocks_for(m_shard, nblk_shard)]
    sc1, ev1, ld1 = full.eigensnp(blk, cfg)
    out["rfit_ev_relerr"] = float(np.abs(ev_r / ev0 - 1).max())
    out["rfit_angle"] = float(pca.subspace_angle(sc_r, sc0))
    out["eigensnp_ev_relerr"] = float(np.abs(ev_e / ev1 - 1).max())
    out["eigensnp_angle"] = float(pca.subspace_angle(sc_e, sc1))
    out["eigensnp_loadings_angle_shard0"] = float(pca.subspace_angle(ld_e, ld1[:m_shard]))
    # (a different test matrix: agreement only to the accuracy of the randomized method itself at k = 8, l = 18 on 21
    #  structural components)
    out["rfit_entropy_seed_ev_relerr"] = float(np.abs(ev_n / ev0 - 1).max())
    out["rfit_entropy_seed_ranks_agree"] = bool(entropy_seed_ranks_agree)
    out["ok"] = bool(out["rfit_entropy_seed_ev_relerr"] < 0.1 and entropy_seed_ranks_agree and out["rfit_ev_relerr"] < 1e-4 and out["rfit_angle"] < 1e-3 and out["eigensnp_ev_relerr"] < 1e-4
                     and out["eigensnp_angle"] < 1e-3)
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
