"""Same-box A/B of the streaming ingest (development aid): e2e = gpca_ingest_bed from pinned host memory + rfit."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, genomic_pca_b200 as gp
n, m = 2504, 10_000_000
dev = torch.device("cuda", 0)
payload = bench.synth_bed_device(torch, n, m, 0, dev)
bps = (n + 3) // 4
host = torch.empty((m, bps), dtype=torch.uint8, pin_memory=True)
host.copy_(payload); torch.cuda.synchronize(); del payload
ctx = gp.Context(0)
out = (np.empty(m, dtype=np.uint8), np.empty(m, dtype=np.float32), np.empty(m, dtype=np.float32))
def run():
    t0 = time.perf_counter()
    ctx.ingest_bed(host.data_ptr(), n, m, qc=None, vcf_maf=0.01, out=out)
    t1 = time.perf_counter()
    ctx.rfit(20, 10, power_iters=2, seed=42, want_loadings=False)
    return (t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3
variants = [{}, {"GPCA_DEBUG_NO_COPY_STREAM": "1"}, {"GPCA_DEBUG_NO_INCR_TRANSPOSE": "1"},
            {"GPCA_DEBUG_NO_COPY_STREAM": "1", "GPCA_DEBUG_NO_INCR_TRANSPOSE": "1"}]
run()
for rnd in range(2):
    for v in variants:
        for k in ("GPCA_DEBUG_NO_COPY_STREAM", "GPCA_DEBUG_NO_INCR_TRANSPOSE"):
            os.environ.pop(k, None)
        os.environ.update(v)
        run()
        r = [run() for _ in range(3)]
        print(sorted(v.keys()) or "default", "ingest ms", [round(x[0], 1) for x in r], "rfit ms", [round(x[1], 1) for x in r], flush=True)
