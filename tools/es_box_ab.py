"""Same-box A/B of the TMA box width of the batched per-LD-block passes (GPCA_I8_ITEM_BOX = 64 / 128) on EigenSNP:
    python tools/es_box_ab.py [samples] [snps] [blocks] [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np          # noqa: E402
import torch                # noqa: E402
import bench                # noqa: E402
import genomic_pca_b200 as gp   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 87_500
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 212
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ctx = gp.Context(0)
cfg = gp.EigenSnpConfig(target_num_global_pcs=20)
ctx.set_memory_reserve(gp.binding.eigensnp_workspace_bytes(n, m, nb, cfg))
host = bench.HostPayload(ctx, n, m, 0)
_, _, _, _, d = ctx.ingest_bed(host.ptr, n, m, qc=gp.QcConfig(0.98, 0.01, 1.0))
edges = np.linspace(0, d, nb + 1).astype(np.int64)
blocks = [np.arange(edges[i], edges[i + 1], dtype=np.uint64) for i in range(nb)]
out = (np.ones((n, 20), dtype=np.float32), np.ones(20), np.ones((d, 20), dtype=np.float32))
res = {}
for rep in range(reps + 1):
    for box in ("64", "128", "default"):
        if box == "default":
            os.environ.pop("GPCA_I8_ITEM_BOX", None)
        else:
            os.environ["GPCA_I8_ITEM_BOX"] = box
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.eigensnp(blocks, cfg, out=out)
        dt = time.perf_counter() - t0
        if rep:
            res.setdefault(box, []).append(dt)
for box, v in res.items():
    print(f"GPCA_I8_ITEM_BOX={box}: eigensnp {np.mean(v) * 1e3:.2f} ms (min {np.min(v) * 1e3:.2f}), resident rows {ctx.resident_snp_rows}/{d}")
host.free()
ctx.close()
