"""Timing breakdown of the host-buffer (e2e) path of bench.py -- development aid."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, genomic_pca_b200 as gp

n, m = 2504, int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
payload = bench.synth_bed_device(torch, n, m, 0, dev)
bps = (n + 3) // 4
host = torch.empty((m, bps), dtype=torch.uint8, pin_memory=True)
host.copy_(payload); torch.cuda.synchronize(); del payload
ctx = gp.Context(0)
for rep in range(2):
    t = [time.perf_counter()]
    ctx.load_bed_host_ptr(host.data_ptr(), n, m); t.append(time.perf_counter())
    keep, mean, sd = ctx.vcf_maf_filter(0.01); t.append(time.perf_counter())
    t.append(time.perf_counter())
    ctx.set_pca_snps_mask(keep, mean, sd); t.append(time.perf_counter())
    ctx.rfit(20, 10, power_iters=2, seed=42, want_loadings=False); t.append(time.perf_counter())
    names = ["load_bed(H2D+repitch)", "vcf_maf_filter(counts+host)", "numpy select", "set_pca_snps", "rfit"]
    print(rep, {k: round((b - a) * 1e3, 1) for k, a, b in zip(names, t[:-1], t[1:])}, "total", round((t[-1] - t[0]) * 1e3, 1))
for rep in range(3):
    t0 = time.perf_counter()
    import ctypes as C
    from genomic_pca_b200 import binding as B
    keep = np.empty(m, dtype=np.uint8); mean = np.empty(m, dtype=np.float32); sd = np.empty(m, dtype=np.float32)
    nn = C.c_uint64(0)
    ta = time.perf_counter()
    rc = B.lib.gpca_ingest_bed(ctx._h, host.data_ptr(), n, m, None, 0, None, 0.01, B._ptr(keep, B._u8p), B._ptr(mean, B._f32p), B._ptr(sd, B._f32p), None, C.byref(nn))
    tb = time.perf_counter()
    print("   raw C call ms", round((tb - ta) * 1e3, 1), "rc", rc)
    t1 = time.perf_counter()
    ctx.rfit(20, 10, power_iters=2, seed=42, want_loadings=False)
    t2 = time.perf_counter()
    print("pipelined", rep, {"ingest_bed": round((t1 - t0) * 1e3, 1), "rfit": round((t2 - t1) * 1e3, 1)}, "total", round((t2 - t0) * 1e3, 1))
