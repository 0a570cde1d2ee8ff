"""Same box, same pinned buffer: plain H2D copy rate against the streaming ingest's (GPCA_TRACE=1 prints its stages)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, genomic_pca_b200 as gp
n, m = 2504, 10_000_000
dev = torch.device("cuda", 0)
payload = bench.synth_bed_device(torch, n, m, 0, dev)
host = torch.empty(payload.shape, dtype=torch.uint8, pin_memory=True)
host.copy_(payload); torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); payload.copy_(host, non_blocking=True); torch.cuda.synchronize()
    print("plain H2D of the payload: %.1f ms = %.1f GB/s" % ((time.perf_counter() - t0) * 1e3, host.numel() / (time.perf_counter() - t0) / 1e9), flush=True)
del payload; torch.cuda.empty_cache()
ctx = gp.Context(0)
out = (np.empty(m, dtype=np.uint8), np.empty(m, dtype=np.float32), np.empty(m, dtype=np.float32))
for rep in range(3):
    t0 = time.perf_counter()
    ctx.ingest_bed(host.data_ptr(), n, m, qc=None, vcf_maf=0.01, out=out)
    print("ingest: %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
