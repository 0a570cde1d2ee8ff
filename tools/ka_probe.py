"""K-a alone: the decode / allele-count kernel on device-resident .bed payloads of the BASELINE shapes.
    python tools/ka_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                # noqa: E402
import bench                # noqa: E402
import genomic_pca_b200 as gp   # noqa: E402

dev = torch.device("cuda", 0)
ctx = gp.Context(0)
for n, m in ((2504, 10_000_000), (500_000, 87_500), (64, 50_000_000), (20_000, 2_000_000)):
    payload = bench.synth_bed_device(torch, n, m, 0, dev)
    bps = (n + 3) // 4
    for grp in [None] + ([int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else []):
        if grp is None:
            os.environ.pop("GPCA_DEBUG_COUNT_GROUP", None)
        else:
            os.environ["GPCA_DEBUG_COUNT_GROUP"] = str(grp)
        ms = ctx.count_kernel_ms(payload.data_ptr(), n, m - 1, reps=10)
        gbs = (m - 1) * bps / ms / 1e6
        print(f"{n} samples x {m} SNPs ({bps} B rows), lanes per row {grp if grp is not None else 'default'}: {ms:.3f} ms per launch, "
              f"{gbs:.0f} GB/s = {gbs / 6544.7:.3f} of the measured HBM peak")
    del payload
    torch.cuda.empty_cache()
ctx.close()
