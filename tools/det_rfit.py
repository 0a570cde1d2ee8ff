import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, genomic_pca_b200 as gp
dev = torch.device("cuda", 0)
n, m = 4096, 100_000
payload = bench.synth_bed_device(torch, n, m, 0, dev)
ctx = gp.Context(0)
ctx.load_bed_device(payload.data_ptr(), n, m)
keep, mean, sd, _ = ctx.snp_qc(gp.QcConfig(0.98, 0.0, 1.0))
ctx.set_pca_snps_mask(keep, mean, sd)
r = [ctx.rfit(10, 10, 2, seed=42) for _ in range(4)]
print({k: v for k, v in os.environ.items() if k.startswith("GPCA_DEBUG")}, [float(np.abs(x[1] - r[0][1]).max()) for x in r[1:]], r[0][1][:3])
