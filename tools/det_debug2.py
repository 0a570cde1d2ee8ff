import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, genomic_pca_b200 as gp
dev = torch.device("cuda", 0)
ENG = int(os.environ.get("ENG", "2"))
def run(n, m):
    payload = bench.synth_bed_device(torch, n, m, 0, dev)
    ctx = gp.Context(0)
    ctx.load_bed_device(payload.data_ptr(), n, m)
    keep, mean, sd, code = ctx.snp_qc(gp.QcConfig(0.98, 0.0, 1.0))
    d = ctx.set_pca_snps_mask(keep, mean, sd)
    ext = torch.cuda.ExternalStream(ctx.stream)
    l = 30
    res = {}
    with torch.cuda.stream(ext):
        g = torch.Generator(device=dev); g.manual_seed(1)
        Bs = torch.randn(n, l, device=dev, generator=g)
        Bd = torch.randn(d, l, device=dev, generator=g)
        ext.synchronize()
        for name, fn, src, rows in (("snp", ctx.sketch_snp_side, Bs, d), ("smp", ctx.sketch_sample_side, Bd, n)):
            ctx.set_sketch_engine(0)
            o = torch.empty(rows, l, device=dev)
            fn(src.data_ptr(), o.data_ptr(), l, l)
            ctx.synchronize()
            ref = o.cpu().numpy()
            ctx.set_sketch_engine(ENG)
            scale = np.abs(ref).max()
            wrong = []
            for rep in range(3):
                o = torch.empty(rows, l, device=dev)
                fn(src.data_ptr(), o.data_ptr(), l, l)
                ctx.synchronize()
                err = np.abs(o.cpu().numpy() - ref).max(axis=1) / scale
                wrong.append(int((err > 2e-3).sum()))
            res[name] = (rows, wrong, float(err.max()))
    del ctx
    return res
for n, m in ((2048, 8000), (2048, 200000), (20000, 8000), (20000, 60000), (600, 400000)):
    print(n, m, {k: v for k, v in os.environ.items() if k.startswith("GPCA_DEBUG")}, run(n, m), flush=True)
