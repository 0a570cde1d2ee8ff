import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, genomic_pca_b200 as gp
dev = torch.device("cuda", 0)
n, m = 2048, 200000
payload = bench.synth_bed_device(torch, n, m, 0, dev)
ctx = gp.Context(0)
ctx.load_bed_device(payload.data_ptr(), n, m)
keep, mean, sd, code = ctx.snp_qc(gp.QcConfig(0.98, 0.0, 1.0))
d = ctx.set_pca_snps_mask(keep, mean, sd)
ext = torch.cuda.ExternalStream(ctx.stream)
l = 30
with torch.cuda.stream(ext):
    g = torch.Generator(device=dev); g.manual_seed(1)
    Bs = torch.randn(n, l, device=dev, generator=g)
    ext.synchronize()
    ctx.set_sketch_engine(0)
    o = torch.empty(d, l, device=dev)
    ctx.sketch_snp_side(Bs.data_ptr(), o.data_ptr(), l, l)
    ctx.synchronize()
    ref = o.cpu().numpy()
    ctx.set_sketch_engine(2)
    o = torch.empty(d, l, device=dev)
    ctx.sketch_snp_side(Bs.data_ptr(), o.data_ptr(), l, l)
    ctx.synchronize()
    out = o.cpu().numpy()
err = np.abs(out - ref).max(axis=1) / np.abs(ref).max()
bad = err > 2e-3
print("rows", d, "bad", int(bad.sum()))
# runs of bad rows
idx = np.nonzero(bad)[0]
runs = []
if len(idx):
    s = idx[0]; p = idx[0]
    for i in idx[1:]:
        if i != p + 1:
            runs.append((s, p - s + 1)); s = i
        p = i
    runs.append((s, p - s + 1))
from collections import Counter
print("run lengths:", Counter(r[1] for r in runs).most_common(10))
print("run start mod 32:", Counter(r[0] % 32 for r in runs).most_common(5), "mod 128:", Counter(r[0] % 128 for r in runs).most_common(8), "mod 256:", Counter(r[0] % 256 for r in runs).most_common(8))
items = Counter(int(r[0] // 256) % 296 for r in runs)
print("CTA (item % 296) with errors:", len(items), "of 296; top", items.most_common(8))
print("item wave (item // 296) histogram:", sorted(Counter(int(r[0] // 256) // 296 for r in runs).items()))
for r in runs[:6]:
    i = r[0]
    print("row", i, "len", r[1], "out", out[i, :4], "ref", ref[i, :4], "ratio-ish", (out[i, :4] - ref[i, :4]))
