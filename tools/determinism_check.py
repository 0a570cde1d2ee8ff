"""Run-to-run bit reproducibility of the sketch passes and the two PCA drivers (development aid).
The integer engine accumulates exactly and every reduction has a fixed order, so repeated calls on the same
context must agree bit for bit."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, genomic_pca_b200 as gp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 60_000
dev = torch.device("cuda", 0)
payload = bench.synth_bed_device(torch, n, m, 0, dev)
ctx = gp.Context(0)
ctx.load_bed_device(payload.data_ptr(), n, m)
keep, mean, sd, code = ctx.snp_qc(gp.QcConfig(0.98, 0.01, 1.0))
d = ctx.set_pca_snps_mask(keep, mean, sd)
out = {"n": n, "d": d}
ext = torch.cuda.ExternalStream(ctx.stream)
l = 30
with torch.cuda.stream(ext):
    g = torch.Generator(device=dev); g.manual_seed(1)
    Bs = torch.randn(n, l, device=dev, generator=g)
    Bd = torch.randn(d, l, device=dev, generator=g)
    ext.synchronize()
    for name, fn, src, rows in (("snp_side", ctx.sketch_snp_side, Bs, d), ("sample_side", ctx.sketch_sample_side, Bd, n)):
        ref = None
        worst = 0.0
        for rep in range(6):
            o = torch.empty(rows, l, device=dev)
            fn(src.data_ptr(), o.data_ptr(), l, l)
            ctx.synchronize()
            if ref is None:
                ref = o
            else:
                worst = max(worst, float((o - ref).abs().max()))
        out[name + "_maxdiff"] = worst
r = [ctx.rfit(20, 10, 2, seed=42) for _ in range(3)]
out["rfit_ev_diff"] = float(max(np.abs(x[1] - r[0][1]).max() for x in r[1:]))
out["rfit_scores_diff"] = float(max(np.abs(x[0] - r[0][0]).max() for x in r[1:]))
nblocks = max(1, d // 400)
edges = np.linspace(0, d, nblocks + 1).astype(np.int64)
blocks = [np.arange(edges[i], edges[i + 1], dtype=np.uint64) for i in range(nblocks)]
cfg = gp.EigenSnpConfig(target_num_global_pcs=20, min_subset_size=2000)
for mode in (True, False):
    ctx.set_batch_blocks(mode)
    e = [ctx.eigensnp(blocks, cfg) for _ in range(3)]
    tag = "eigensnp_batched" if mode else "eigensnp_per_block"
    out[tag + "_ev_diff"] = float(max(np.abs(x[1] - e[0][1]).max() for x in e[1:]))
    out[tag + "_scores_diff"] = float(max(np.abs(x[0] - e[0][0]).max() for x in e[1:]))
    out[tag + "_ev0"] = float(e[0][1][0])
print(json.dumps(out))
