import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, genomic_pca_b200 as gp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 60_000
eng = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
payload = bench.synth_bed_device(torch, n, m, 0, dev)
ctx = gp.Context(0)
ctx.set_sketch_engine(eng)
ctx.load_bed_device(payload.data_ptr(), n, m)
keep, mean, sd, code = ctx.snp_qc(gp.QcConfig(0.98, 0.01, 1.0))
d = ctx.set_pca_snps_mask(keep, mean, sd)
ext = torch.cuda.ExternalStream(ctx.stream)
l = 30
with torch.cuda.stream(ext):
    g = torch.Generator(device=dev); g.manual_seed(1)
    Bs = torch.randn(n, l, device=dev, generator=g)
    Bd = torch.randn(d, l, device=dev, generator=g)
    ext.synchronize()
    for name, fn, src, rows in (("snp_side", ctx.sketch_snp_side, Bs, d), ("sample_side", ctx.sketch_sample_side, Bd, n)):
        outs = []
        for rep in range(6):
            o = torch.empty(rows, l, device=dev)
            fn(src.data_ptr(), o.data_ptr(), l, l)
            ctx.synchronize()
            torch.cuda.synchronize()
            outs.append(o.cpu().numpy())
        for rep in range(1, 6):
            diff = np.abs(outs[rep] - outs[0])
            bad = np.argwhere(diff > 0)
            rows_bad = np.unique(bad[:, 0])
            print(name, "rep", rep, "n_diff", len(bad), "rows", len(rows_bad), "first rows", rows_bad[:8].tolist(), "last", rows_bad[-3:].tolist() if len(rows_bad) else [],
                  "cols", np.unique(bad[:, 1])[:40].tolist(), "max", float(diff.max()))
        # reference for the first 2048 output rows in torch f64
        if name == "snp_side":
            pass
