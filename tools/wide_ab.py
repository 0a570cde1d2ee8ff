"""Same-box A/B of a run-time switch of the integer engine -- by default its two CTA shapes (GPCA_I8_WIDE: 256-row CTAs,
two per SM / 512-row CTAs, one per SM); GPCA_I8_TILE_SYNC selects the per-row-tile TMEM hand-over:
    python tools/wide_ab.py [samples] [snps] [reps] [ENV_NAME | NAME=V[,NAME=V] ...]
(one ENV_NAME: 0 against 1; several NAME=V specs: the default build against each of them)
rfit (k = 20, l = 30, q = 2) on a device-generated matrix; the sketch kernel's own launch times (CUDA events) per pass
orientation, alternating the two shapes rep by rep so that clock / power drift hits both alike."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np          # noqa: E402
import torch                # noqa: E402
import bench                # noqa: E402
import genomic_pca_b200 as gp   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 87_500
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
specs = sys.argv[4:] if len(sys.argv) > 4 else ["GPCA_I8_WIDE"]
if len(specs) == 1 and "=" not in specs[0]:
    variants = {specs[0] + "=0": {specs[0]: "0"}, specs[0] + "=1": {specs[0]: "1"}}
else:
    variants = {"default": {}}
    for sp in specs:
        variants[sp] = dict(kv.split("=") for kv in sp.split(","))
all_names = sorted({k for v in variants.values() for k in v})
dev = torch.device("cuda", 0)
ctx = gp.Context(0)
ctx.set_sketch_timing(True)
host = bench.HostPayload(ctx, n, m, 0)
ctx.ingest_bed(host.ptr, n, m, qc=None, vcf_maf=0.01)
out = (np.ones((n, 20)), np.ones(20), None)
res = {}
for rep in range(reps + 1):
    for wide, envs in variants.items():
        for k in all_names:
            os.environ.pop(k, None)
        os.environ.update(envs)
        ctx.sketch_stats(reset=True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.ExternalStream(ctx.stream, device=dev)
        ev0.record(st)
        for _ in range(3):
            ctx.rfit(20, 10, power_iters=2, seed=42, want_loadings=False, out=out)
        ev1.record(st)
        torch.cuda.synchronize()
        ms, _, npass = ctx.sketch_stats(reset=True)
        if rep:      # first round = warm-up
            res.setdefault(wide, []).append((ctx.last_kernel_ms / npass, ev0.elapsed_time(ev1) / 3))
for wide, v in res.items():
    k = np.array([x[0] for x in v])
    s = np.array([x[1] for x in v])
    gb = ctx.num_pca_snps * ((n + 3) // 4) / 1e9
    print(f"{wide}: kernel {k.mean():.3f} ms/launch (min {k.min():.3f})  {gb / k.mean() * 1e3:.0f} GB/s "
          f"= {gb / k.mean() * 1e3 / 6544.7:.3f} of HBM peak; rfit step {s.mean():.2f} ms")
host.free()
ctx.close()
