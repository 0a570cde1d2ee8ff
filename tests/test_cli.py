"""Drop-in CLI (genomic_pca_b200/genomic_pca): flag surface / messages on the CPU, full workflows on the GPU."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from oracle import bed, vcf

from conftest import ROOT
from helpers import make_dataset

CLI = os.path.join(ROOT, "genomic_pca_b200", "genomic_pca")


def _run(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True)


def test_cli_argument_surface():
    assert os.path.exists(CLI), "build the CLI with __graft_entry__.build()"
    h = _run("--help")
    for flag in ["--vcf-dir", "--components", "--maf", "--rfit-seed", "--eigensnp", "--bed-file", "--ld-block-file",
                 "--eigensnp-sample-keep-file", "--eigensnp-min-call-rate", "--eigensnp-min-maf", "--eigensnp-max-hwe-p",
                 "--eigensnp-k-global", "--eigensnp-components-per-block", "--eigensnp-subset-factor",
                 "--eigensnp-min-subset-size", "--eigensnp-max-subset-size", "--eigensnp-global-oversampling",
                 "--eigensnp-global-power-iter", "--eigensnp-local-oversampling", "--eigensnp-local-power-iter",
                 "--eigensnp-seed", "--eigensnp-snp-strip-size", "--eigensnp-refine-passes",
                 "--eigensnp-collect-diagnostics", "--out", "--threads", "--log-level"]:
        assert flag in h.stderr, flag                                      # src/main.rs:505-592
    r = _run("-o", "/tmp/x")
    assert r.returncode != 0 and "--vcf-dir is required" in r.stderr      # main.rs:116
    r = _run("-o", "/tmp/x", "-d", "/tmp")
    assert r.returncode != 0 and "--components is required" in r.stderr   # main.rs:119
    r = _run("-o", "/tmp/x", "--eigensnp")
    assert r.returncode != 0 and "--bed-file is required" in r.stderr     # main.rs:296
    r = _run("-d", "/tmp", "-k", "3")
    assert r.returncode != 0 and "--out" in r.stderr


# value -> what Rust's `format!("{:.6}", v)` prints (src/main.rs:721,755,779,832): exact decimal expansion of the binary
# value, round-half-to-even at the sixth place, sign of negative zero kept, "NaN" / "inf" / "-inf"
RUST_F6 = [
    ("0", "0.000000"), ("-0.0", "-0.000000"), ("1.5", "1.500000"), ("-2.25", "-2.250000"),
    ("2.5e-7", "0.000000"), ("5e-7", "0.000000"), ("1.5e-6", "0.000002"), ("-1e-7", "-0.000000"),
    ("0.0078125", "0.007812"),          # exact tie (2^-7): half-to-even keeps the even digit
    ("0.0234375", "0.023438"),          # exact tie, odd digit rounds up
    ("0.9999995", "1.000000"),          # the f64 nearest to it is 0.99999950000000002...: above the tie
    ("0.9999994", "0.999999"),
    ("123456.7890125", "123456.789012"),
    ("1e21", "1000000000000000000000.000000"),
    ("nan", "NaN"), ("-nan", "NaN"), ("inf", "inf"), ("-inf", "-inf"),
]


def test_writers_number_format_is_rusts():
    """The CLI's and the Python writers' `{:.6}`: byte for byte what the reference's writers print, including the edge
    values (ties, negative zero, infinities, NaN) -- and for f32 inputs (EigenSNP scores / loadings are f32 widened to
    their exact value, as Rust prints an f32)."""
    from genomic_pca_b200 import plink
    r = _run("--format-f6", *[v for v, _ in RUST_F6], "0.1", "16777217")
    assert r.returncode == 0, r.stderr
    rows = [ln.split("\t") for ln in r.stdout.splitlines()]
    assert [row[0] for row in rows[:len(RUST_F6)]] == [want for _, want in RUST_F6]
    assert rows[-2] == ["0.100000", "0.100000"] and rows[-1] == ["16777217.000000", "16777216.000000"]
    for v, want in RUST_F6:
        assert plink._fmt6(float(v)) == want, v
    # an independent statement of the rule for random values: exact decimal expansion, half-to-even at 1e-6
    from decimal import Decimal, ROUND_HALF_EVEN
    rng = np.random.default_rng(5)
    vals = np.concatenate([rng.standard_normal(200) * 10.0 ** rng.integers(-8, 6, 200),
                           (rng.integers(-10**6, 10**6, 50) * 2 + 1) / 2.0 ** 21])       # exact ties among them
    r = _run("--format-f6", *[repr(float(v)) for v in vals])
    got = [ln.split("\t")[0] for ln in r.stdout.splitlines()]
    for v, g in zip(vals, got):
        want = str(Decimal(float(v)).quantize(Decimal("0.000001"), rounding=ROUND_HALF_EVEN))
        if float(v) < 0 and want.lstrip("-") == "0.000000":
            want = "-0.000000"
        assert g == want == plink._fmt6(float(v)), (v, g, want)
    assert plink._fmt6(float(np.float32(0.1))) == "0.100000"
    # exact expansion of an f32 that is not an f64-rounded decimal: 0.3f = 0.300000011920928955078125
    assert plink._fmt6(float(np.float32(0.3)) * 1e0) == "0.300000"


def _write_plink(tmp, g, payload, chrom="1"):
    m, n = g.shape
    (tmp / "d.bed").write_bytes(bytes([0x6C, 0x1B, 0x01]) + payload.tobytes())
    (tmp / "d.fam").write_text("".join(f"F{i} S{i} 0 0 0 -9\n" for i in range(n)))
    (tmp / "d.bim").write_text("".join(f"{chrom}\trs{j}\t0\t{1000 + 10 * j}\tA\tG\n" for j in range(m)))


@pytest.mark.gpu
def test_cli_eigensnp_workflow(tmp_path, gpu_ctx):
    import genomic_pca_b200 as gp
    from genomic_pca_b200 import plink
    g, payload = make_dataset(400, 1500, n_pops=4, seed=31)
    _write_plink(tmp_path, g, payload)
    (tmp_path / "ld.txt").write_text("# blocks\nchr1 1000 5990\n1 6000 10990\nchr1 11000 99999999\n")
    out = tmp_path / "res" / "run"
    r = _run("--eigensnp", "--bed-file", str(tmp_path / "d.bed"), "--ld-block-file", str(tmp_path / "ld.txt"),
             "-o", str(out), "--eigensnp-k-global", "3", "--eigensnp-min-subset-size", "100",
             "--eigensnp-max-subset-size", "300", "--eigensnp-subset-factor", "0.5", "--eigensnp-seed", "9",
             "--eigensnp-collect-diagnostics")
    assert r.returncode == 0, r.stderr
    import json
    diag = json.loads((tmp_path / "res" / "run.eigensnp_diagnostics.json").read_text())      # main.rs:411-430
    assert diag["num_qc_samples"] == 400 and diag["num_ld_blocks"] == 3 and diag["components"] == 3
    assert [s["stage"] for s in diag["stages_ms"]][-1] == "outputs" and diag["kernel_launches"] > 0
    pcs = (tmp_path / "res" / "run.eigensnp.pca.tsv").read_text().splitlines()
    assert pcs[0] == "SampleID\tPC1\tPC2\tPC3" and len(pcs) == 401 and pcs[1].split("\t")[0] == "S0"
    evs = (tmp_path / "res" / "run.eigenvalues.tsv").read_text().splitlines()
    assert evs[0] == "PC\tEigenvalue" and len(evs) == 4
    lds = (tmp_path / "res" / "run.eigensnp.loadings.tsv").read_text().splitlines()
    assert lds[0] == "VariantID\tChrom\tPos\tPC1_loading\tPC2_loading\tPC3_loading"
    # same numbers as the library driven from Python with the same configuration
    chrom, sid, bp = plink.read_bim(str(tmp_path / "d.bim"))
    prep = plink.prepare_data_for_eigen_snp(gpu_ctx, payload, plink.read_fam(str(tmp_path / "d.fam")), chrom, bp,
                                            plink.parse_ld_block_file(str(tmp_path / "ld.txt")))
    cfg = gp.EigenSnpConfig(target_num_global_pcs=3, min_subset_size=100, max_subset_size=300, subset_factor=0.5,
                            random_seed=9)
    sc, ev, load = gpu_ctx.eigensnp(prep["block_snp_ids"], cfg)
    got = np.array([[float(x) for x in ln.split("\t")[1:]] for ln in pcs[1:]])
    assert np.abs(got - sc).max() <= 1e-6 + 1e-6 * np.abs(sc).max()
    got_ev = np.array([float(ln.split("\t")[1]) for ln in evs[1:]])
    assert np.abs(got_ev - ev).max() <= 1e-6 + 1e-6 * np.abs(ev).max()
    assert len(lds) - 1 == prep["pca_original_idx"].size
    assert lds[1].split("\t")[0] == f"rs{prep['pca_original_idx'][0]}"


@pytest.mark.gpu
def test_cli_vcf_workflow(tmp_path, gpu_ctx):
    g, _ = make_dataset(120, 900, n_pops=3, seed=33)
    g = g.astype(np.uint8)

    def vcf_text(rows, chrom, with_bad):
        lines = ["##fileformat=VCFv4.2", '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
                 "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{i}" for i in range(120))]
        gt = {0: "0|0", 1: "0|1", 2: "1/1"}
        for j, row in enumerate(rows):
            lines.append(f"{chrom}\t{100 + j}\t.\tA\tC\t.\t.\t.\tGT:DP\t" + "\t".join(gt[int(v)] + ":7" for v in row))
        if with_bad:
            lines.append(f"{chrom}\t5000\t.\tAT\tC\t.\t.\t.\tGT\t" + "\t".join(["0|1"] * 120))          # REF len 2
            lines.append(f"{chrom}\t5001\t.\tA\tC,G\t.\t.\t.\tGT\t" + "\t".join(["0|1"] * 120))         # multi-allelic
            lines.append(f"{chrom}\t5002\t.\tA\tC\t.\t.\t.\tGT\t" + "\t".join(["./."] + ["0|1"] * 119))  # missing
        return "\n".join(lines) + "\n"

    d = tmp_path / "vcfs"
    d.mkdir()
    t2 = vcf_text(g[500:], "2", True)
    with gzip.open(d / "b.chr2.vcf.gz", "wt") as f:
        f.write(t2)
    t1 = vcf_text(g[:500], "1", False)
    (d / "a.chr1.vcf").write_text(t1)
    (d / "notes.txt").write_text("ignored")
    out = tmp_path / "o" / "v"
    r = _run("-d", str(d), "-o", str(out), "-k", "3", "--maf", "0.05", "--rfit-seed", "11")
    assert r.returncode == 0, r.stderr
    pcs = (tmp_path / "o" / "v.vcf.pca.tsv").read_text().splitlines()
    assert pcs[0] == "SampleID\tPC1\tPC2\tPC3" and len(pcs) == 121 and pcs[1].startswith("s0\t")
    assert (tmp_path / "o" / "v.eigenvalues.tsv").read_text() == "PC\tEigenvalue\n"        # main.rs:676 quirk kept
    # same numbers as: oracle VCF filter semantics -> library rfit with the same seed
    s1, ids1, d1 = vcf.parse_vcf_text(t1, 0.05)
    s2, ids2, d2 = vcf.parse_vcf_text(t2, 0.05)
    dos = np.concatenate([d1, d2])                     # files in sorted-path order (main.rs:152)
    gpu_ctx.load_u8_variant_major(dos)
    keep, mean, sd = gpu_ctx.vcf_maf_filter(0.05)
    assert keep.all()
    gpu_ctx.set_pca_snps_mask(keep, mean, sd)
    sc, ev, _ = gpu_ctx.rfit(3, 10, power_iters=2, seed=11, want_loadings=False)
    got = np.array([[float(x) for x in ln.split("\t")[1:]] for ln in pcs[1:]])
    assert np.abs(got - sc).max() <= 1e-6 + 1e-6 * np.abs(sc).max()


# ------------------------------------------------------------------------------------------- VCF text on the host
def _bgzf(data: bytes, block: int = 0xff00) -> bytes:
    """BGZF as bgzip writes it: gzip members of <= 64 KiB input with the 'BC' extra field, then the empty EOF member."""
    import struct
    import zlib

    def member(chunk):
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        raw = c.compress(chunk) + c.flush()
        head = b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, len(raw) + 25)
        return head + raw + struct.pack("<II", zlib.crc32(chunk), len(chunk))

    out = b"".join(member(data[i:i + block]) for i in range(0, len(data), block))
    return out + member(b"")


def _vcf_text(g, chrom, n, fmt="GT:DP", extra=(), crlf=False, final_newline=True):
    lines = ["##fileformat=VCFv4.2", '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
             "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{i}" for i in range(n))]
    gt = {0: "0|0", 1: "0|1", 2: "1/1"}
    gi = fmt.split(":").index("GT")
    for j, row in enumerate(g):
        def cell(v):
            parts = ["7"] * len(fmt.split(":"))
            parts[gi] = gt[int(v)]
            return ":".join(parts)
        lines.append(f"{chrom}\t{100 + j}\trs{j}\tA\tC\t.\t.\t.\t{fmt}\t" + "\t".join(cell(v) for v in row))
    lines += list(extra)
    t = ("\r\n" if crlf else "\n").join(lines)
    return t + (("\r\n" if crlf else "\n") if final_newline else "")


def _pack_plink(dos):
    """dosage u8 [D x N] -> 2-bit rows in PLINK coding (0 -> 11, 1 -> 10, 2 -> 00), zero padding"""
    d, n = dos.shape
    code = np.array([3, 2, 0], dtype=np.uint8)[dos]
    pad = (-n) % 4
    code = np.pad(code, ((0, 0), (0, pad)))
    c4 = code.reshape(d, -1, 4)
    return (c4[:, :, 0] | (c4[:, :, 1] << 2) | (c4[:, :, 2] << 4) | (c4[:, :, 3] << 6)).astype(np.uint8)


@pytest.mark.parametrize("threads,batch", [(1, 1 << 26), (4, 1 << 26), (8, 3000), (3, 150_000)])
def test_vcf_host_parser_plain_gzip_bgzf(tmp_path, threads, batch):
    """The host half of the VCF workflow without a GPU (`--parse-vcf`): three containers (plain text, one gzip stream,
    BGZF with its blocks inflated concurrently), batches smaller than a line, CRLF, a last line without a newline,
    GT at another FORMAT position, the reference's drop rules -- ids and 2-bit rows byte for byte what the oracle's
    restatement of src/vcf.rs gives, in sorted-path order."""
    import json
    n = 203                                                # not a multiple of 4: padded last byte
    g, _ = make_dataset(n, 1500, n_pops=3, seed=91)
    g = g.astype(np.uint8)
    bad = [f"3\t9000\t.\tAT\tC\t.\t.\t.\tGT\t" + "\t".join(["0|1"] * n),            # REF longer than one base
           f"3\t9001\t.\tA\tC,G\t.\t.\t.\tGT\t" + "\t".join(["0|1"] * n),           # multi-allelic
           f"3\t9002\t.\tA\t.\t.\t.\t.\tGT\t" + "\t".join(["0|1"] * n),             # no ALT
           f"3\t9003\t.\tA\tC\t.\t.\t.\tGT\t" + "\t".join(["./."] + ["0|1"] * (n - 1)),   # a missing call
           f"3\t9004\t.\tA\tC\t.\t.\t.\tDP\t" + "\t".join(["7"] * n),               # no GT key
           f"3\t9005\t.\tA\tC\t.\t.\t.\tGT\t" + "\t".join(["0|1"] * (n - 1)),        # a sample short
           f"3\t9006\t.\tA\tC\t.\t.\t.\tGT\t" + "\t".join(["0|2"] + ["0|1"] * (n - 1)),   # allele other than 0 / 1
           f"3\t9007\t.\tA\tC\t.\t.\t.\tGT\t" + "\t".join(["0|0"] * n)]              # monomorphic: MAF filter
    texts = {
        "a.chr1.vcf": _vcf_text(g[:400], "1", n, final_newline=False),
        "b.chr2.vcf.gz": _vcf_text(g[400:900], "2", n, fmt="DP:GT:GQ", crlf=True),
        "c.chr3.vcf.gz": _vcf_text(g[900:], "3", n, extra=bad),
    }
    d = tmp_path / "vcfs"
    d.mkdir()
    (d / "a.chr1.vcf").write_text(texts["a.chr1.vcf"])
    with gzip.open(d / "b.chr2.vcf.gz", "wb") as f:
        f.write(texts["b.chr2.vcf.gz"].encode())
    (d / "c.chr3.vcf.gz").write_bytes(_bgzf(texts["c.chr3.vcf.gz"].encode(), block=20_000))
    (d / "notes.txt").write_text("ignored")
    r = _run("--parse-vcf", str(d), "0.05", str(threads), str(batch), str(tmp_path / "dump"))
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout)
    ids, rows = [], []
    for name in sorted(texts):
        s, i, dos = vcf.parse_vcf_text(texts[name].replace("\r\n", "\n"), 0.05)
        assert len(s) == n
        ids += i
        rows.append(dos)
    dos = np.concatenate(rows)
    assert info["samples"] == n and info["variants"] == len(ids) and info["files"] == 3
    assert info["bgzf_blocks"] >= len(texts["c.chr3.vcf.gz"]) // 20_000          # the BGZF path really ran
    assert (tmp_path / "dump.ids").read_text().split("\n")[:-1] == ids
    got = np.frombuffer((tmp_path / "dump.packed").read_bytes(), dtype=np.uint8).reshape(len(ids), -1)
    assert np.array_equal(got, _pack_plink(dos))


def test_vcf_host_parser_errors(tmp_path):
    n = 8
    g = np.random.default_rng(3).integers(0, 3, size=(50, n)).astype(np.uint8)
    d = tmp_path / "v"
    d.mkdir()
    blob = bytearray(_bgzf(_vcf_text(g, "1", n).encode(), block=700))
    blob[len(blob) // 2] ^= 0x55                                     # a flipped byte inside a block
    (d / "x.vcf.gz").write_bytes(bytes(blob))
    r = _run("--parse-vcf", str(d), "0.0", "2", "100000")
    assert r.returncode != 0 and ("BGZF" in r.stderr)
    (d / "x.vcf.gz").unlink()
    (d / "a.vcf").write_text(_vcf_text(g, "1", n))
    (d / "b.vcf").write_text(_vcf_text(g, "2", n + 1))                # another sample set
    r = _run("--parse-vcf", str(d), "0.0", "2", "100000")
    assert r.returncode != 0 and "Sample mismatch" in r.stderr       # main.rs:157-160
    (d / "b.vcf").write_text(_vcf_text(g, "2", n).replace('##FORMAT=<ID=GT', '##FORMAT=<ID=XX'))
    r = _run("--parse-vcf", str(d), "0.0", "2", "100000")
    assert r.returncode != 0 and "GT key" in r.stderr                # vcf.rs:93
