"""genomic_pca --gpus 2 against --gpus 1 on the same files (EigenSNP and VCF workflows): same retained SNPs, eigenvalues
within 1e-4, scores / loadings up to sign within the north-star tolerances.  Needs 2 GPUs:
    python tools/cli_multi_gpu_check.py [workdir]"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_dataset      # noqa: E402
from oracle import pca                # noqa: E402

CLI = os.path.join(ROOT, "genomic_pca_b200", "genomic_pca")
work = sys.argv[1] if len(sys.argv) > 1 else tempfile.mkdtemp()
os.makedirs(work, exist_ok=True)
n, m = 900, 6000
g, payload = make_dataset(n, m, n_pops=5, seed=2027)
open(os.path.join(work, "d.bed"), "wb").write(bytes([0x6C, 0x1B, 0x01]) + payload.tobytes())
open(os.path.join(work, "d.fam"), "w").write("".join(f"F{i} S{i} 0 0 0 -9\n" for i in range(n)))
open(os.path.join(work, "d.bim"), "w").write("".join(f"1\trs{j}\t0\t{1000 + 10 * j}\tA\tG\n" for j in range(m)))
# 14 blocks of ~400 SNPs, a gap (SNPs in no block), tag-sorted order != genomic order
lines = []
for b in range(14):
    lo = 1000 + 10 * (b * 420)
    lines.append(f"chr1 {lo} {lo + 10 * 399}\n")
open(os.path.join(work, "ld.txt"), "w").write("".join(lines))


def run(tag, gpus, extra):
    out = os.path.join(work, tag)
    r = subprocess.run([CLI, "-o", out, "--gpus", str(gpus), *extra], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out, r.stderr


def table(path, skip_cols=1):
    rows = open(path).read().splitlines()[1:]
    return np.array([[float(x) for x in ln.split("\t")[skip_cols:]] for ln in rows]), [ln.split("\t")[0] for ln in rows]


res = {}
es = ["--eigensnp", "--bed-file", os.path.join(work, "d.bed"), "--ld-block-file", os.path.join(work, "ld.txt"),
      "--eigensnp-k-global", "4", "--eigensnp-min-subset-size", "300", "--eigensnp-max-subset-size", "600",
      "--eigensnp-subset-factor", "0.5", "--eigensnp-seed", "9"]
o1, _ = run("es1", 1, es)
o2, log2 = run("es2", 2, es)
sc1, ids1 = table(o1 + ".eigensnp.pca.tsv")
sc2, ids2 = table(o2 + ".eigensnp.pca.tsv")
ev1, _ = table(o1 + ".eigenvalues.tsv")
ev2, _ = table(o2 + ".eigenvalues.tsv")
ld1, v1 = table(o1 + ".eigensnp.loadings.tsv", 3)
ld2, v2 = table(o2 + ".eigensnp.loadings.tsv", 3)
res["eigensnp"] = {"same_samples": ids1 == ids2, "same_variants": v1 == v2, "n_variants": len(v1),
                   "ev_relerr": float(np.abs(ev2 / ev1 - 1).max()), "score_angle": pca.subspace_angle(sc1, sc2),
                   "loading_angle": pca.subspace_angle(ld1, ld2), "sharded_log": "Sharding over 2 GPUs" in log2}
vdir = os.path.join(work, "vcfs")
os.makedirs(vdir, exist_ok=True)
gt = {0: "0|0", 1: "0|1", 2: "1/1"}
with open(os.path.join(vdir, "a.chr1.vcf"), "w") as f:
    f.write("##fileformat=VCFv4.2\n##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n")
    f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{i}" for i in range(n)) + "\n")
    for j in range(3000):
        f.write(f"1\t{100 + j}\t.\tA\tC\t.\t.\t.\tGT\t" + "\t".join(gt[int(v)] for v in g[j]) + "\n")
vc = ["-d", vdir, "-k", "4", "--rfit-seed", "11", "--write-eigenvalues"]
o1, _ = run("v1", 1, vc)
o2, _ = run("v2", 2, vc)
sc1, _ = table(o1 + ".vcf.pca.tsv")
sc2, _ = table(o2 + ".vcf.pca.tsv")
ev1, _ = table(o1 + ".eigenvalues.tsv")
ev2, _ = table(o2 + ".eigenvalues.tsv")
res["vcf"] = {"ev_relerr": float(np.abs(ev2 / ev1 - 1).max()), "score_angle": pca.subspace_angle(sc1, sc2)}
res["ok"] = bool(res["eigensnp"]["same_variants"] and res["eigensnp"]["sharded_log"] and res["eigensnp"]["ev_relerr"] < 1e-4
                 and res["eigensnp"]["score_angle"] < 1e-3 and res["eigensnp"]["loading_angle"] < 1e-3
                 and res["vcf"]["ev_relerr"] < 1e-4 and res["vcf"]["score_angle"] < 1e-3)
print(json.dumps(res))
sys.exit(0 if res["ok"] else 1)
