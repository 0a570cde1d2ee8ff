#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's own data and importable Python helpers.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py

What comes from the REFERENCE itself (not from oracle/):
  * ``decode_ref``  -- tests/disk.py:89-137 ``process_snp_block_for_accessor_test`` run on
    rows of data/chr22_subset50.bed (its convention: 00->0, 10->1, 11->2, 01->255).
  * ``hwe_ref``     -- tests/pca.py:54-66 ``hwe_pval`` (an independent HWE chi-square in the
    reference repository; ``bed_reader`` is stubbed because only that function is used).
  * ``keep_ref``    -- the QC ladder of tests/pca.py:86-105 (call rate, MAF, HWE, variance), executed FROM THE
    REFERENCE'S OWN SOURCE TEXT on the whole fixture, 2000 variants per batch as its ``--variant-chunk`` default, with
    ``bed_reader.open_bed`` replaced by a decoder of the same bytes (count_A1=False: 00->0, 10->1, 11->2, 01->NaN).
    Stored with the whole payload (17 MB of 2-bit rows compress to ~1 MB) in chr22_subset50_full.npz, together with
    the variants on which that float32 script and the Rust ladder (src/prepare.rs:1283-1364, restated in oracle/bed.py)
    disagree, each with the reason.
What is stored beside them: the packed rows they were computed from, so the tests can feed
the same bytes to oracle/ and to the CUDA path anywhere.
"""
import argparse
import importlib.util
import inspect
import io
import json
import os
import sys
import types
import zipfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(path, name, stubs=()):
    for s in stubs:
        m = types.ModuleType(s)
        m.open_bed = None
        sys.modules.setdefault(s, m)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    disk = _load(f"{REF}/tests/disk.py", "ref_disk")
    refpca = _load(f"{REF}/tests/pca.py", "ref_pca", stubs=("bed_reader",))

    bed_bytes = zipfile.ZipFile(f"{REF}/data/chr22_subset50.bed.zip").read("chr22_subset50.bed")
    fam_txt = zipfile.ZipFile(f"{REF}/data/chr22_subset50.fam.zip").read("chr22_subset50.fam").decode()
    iids = [ln.split()[1] for ln in fam_txt.splitlines() if ln.strip()]
    n = len(iids)
    bps = (n + 3) // 4
    m = (len(bed_bytes) - 3) // bps
    assert bed_bytes[:3] == bytes([0x6C, 0x1B, 0x01])

    # rows: a contiguous head (for block/strip tests) + a strided sample of the whole file
    head = np.arange(0, 4096)
    strided = np.arange(4096, m, 2081)[:1024]
    rows = np.concatenate([head, strided])
    payload = np.frombuffer(bed_bytes, dtype=np.uint8, offset=3).reshape(m, bps)[rows].copy()

    # (1) reference decode of those rows through tests/disk.py
    sample_idx = list(range(n))
    dec = disk.process_snp_block_for_accessor_test(bed_bytes, [int(r) for r in rows], sample_idx, n)
    decode_ref = np.array(dec, dtype=np.uint8)                      # [rows, N], 255 = missing

    # (2) reference HWE p-values through tests/pca.py on the observed count triples + edge cases
    a0 = (decode_ref == 0).sum(1)
    a1 = (decode_ref == 1).sum(1)
    a2 = (decode_ref == 2).sum(1)
    trip = np.stack([a0, a1, a2], 1)
    extra = np.array([[10, 0, 10], [0, 20, 0], [1, 0, 63], [30, 30, 4], [0, 0, 64], [64, 0, 0],
                      [5000, 200, 40], [100000, 50000, 7000], [12, 40, 12], [0, 1, 63],
                      [250000, 200000, 50000], [400, 90, 10]], dtype=np.int64)
    trip = np.concatenate([np.unique(trip, axis=0), extra])
    hwe_ref = np.array([refpca.hwe_pval(int(a), int(b), int(c)) for a, b, c in trip], dtype=np.float64)

    np.savez_compressed(os.path.join(HERE, "chr22_subset50_rows.npz"),
                        n_samples=np.int64(n), n_snps_total=np.int64(m), rows=rows.astype(np.int64),
                        payload=payload, decode_ref=decode_ref, iids=np.array(iids),
                        hwe_triples=trip.astype(np.int64), hwe_ref=hwe_ref)

    # (3) whole-file summary (needs the full .bed; checked on CPU when the reference is mounted)
    codes = np.zeros(4, dtype=np.int64)
    full = np.frombuffer(bed_bytes, dtype=np.uint8, offset=3)
    for j in range(4):
        codes += np.bincount((full >> (2 * j)) & 3, minlength=4)
    json.dump(dict(n_samples=n, n_snps=m, bytes_per_snp=bps, code_hist=[int(c) for c in codes],
                   note="code_hist counts 2-bit fields of all payload bytes (pad bits included: N%4==0 here)"),
              open(os.path.join(HERE, "chr22_subset50_summary.json"), "w"), indent=1)
    print("rows", rows.shape, "payload", payload.shape, "hwe triples", trip.shape)

    # (4) the reference's own QC ladder (tests/pca.py:86-105) on the whole fixture
    full_payload = np.frombuffer(bed_bytes, dtype=np.uint8, offset=3).reshape(m, bps)
    keep_ref, parts = run_reference_ladder(refpca, full_payload, n)
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import bed as obed
    dos = obed.decode_count_a1(full_payload, n)
    nv, n0, n1, n2, _ = obed.snp_counts(dos)
    keep_rust, mean, sd, code = obed.qc_from_counts(n, nv, n0, n1, n2)
    diff = np.nonzero(keep_ref != keep_rust)[0]
    reasons = []
    for j in diff:
        why = []
        if parts["hwe"][j] > 1e-6 and code[j] == 5:
            why.append(f"HWE p: float32 {parts['hwe'][j]!r} > 1e-6 in pca.py, f64 <= 1e-6 in the Rust ladder")
        if parts["hwe"][j] <= 1e-6 and code[j] != 5 and keep_rust[j]:
            why.append(f"HWE p: pca.py's own chi-square gives float32 {parts['hwe'][j]!r} <= 1e-6, the Rust ladder's > 1e-6")
        if not (parts["var"][j] > 1e-9) and keep_rust[j]:
            why.append("variance: float32 nanvar in pca.py")
        if not why:
            why.append(f"pca.py call_rate={parts['call_rate'][j]!r} maf={parts['maf'][j]!r} hwe={parts['hwe'][j]!r} "
                       f"var={parts['var'][j]!r}; Rust-ladder fail code {int(code[j])}")
        reasons.append("; ".join(why))
    np.savez_compressed(os.path.join(HERE, "chr22_subset50_full.npz"), n_samples=np.int64(n), payload=full_payload,
                        keep_ref_bits=np.packbits(keep_ref), keep_rust_bits=np.packbits(keep_rust),
                        diff_idx=diff.astype(np.int64), diff_reason=np.array(reasons))
    print("whole fixture: pca.py ladder keeps", int(keep_ref.sum()), "; Rust ladder keeps", int(keep_rust.sum()),
          "; they differ on", diff.size, "variants")
    for j, r in zip(diff[:20], reasons[:20]):
        print("  variant", int(j), ":", r)


def run_reference_ladder(refpca, payload, n):
    """Execute the statements of tests/pca.py's batch loop from `# QC metrics` to the end of the `ok = ...` expression
    (source lines taken from the reference at generation time, not copied here) on every 2000-variant batch."""
    src = inspect.getsource(refpca.main).splitlines()
    i0 = next(i for i, ln in enumerate(src) if "# QC metrics" in ln)
    i1 = next(i for i, ln in enumerate(src) if "if ok.any():" in ln)
    body = "\n".join(ln[8:] if ln.startswith("        ") else ln.lstrip() for ln in src[i0:i1])
    code = compile(body, "/root/reference/tests/pca.py[QC ladder]", "exec")
    args = argparse.Namespace(min_call_rate=0.98, min_maf=0.01, max_hwe_p=1e-6, min_variance_epsilon=1e-9)  # cli() defaults
    lut = np.array([0.0, np.nan, 1.0, 2.0], dtype=np.float32)       # count_A1=False (tests/pca.py:74)
    m = payload.shape[0]
    keep = np.zeros(m, dtype=bool)
    parts = {k: np.zeros(m, dtype=np.float64) for k in ("call_rate", "maf", "hwe", "var")}
    shifts = np.array([0, 2, 4, 6], dtype=np.uint8)
    import warnings
    for start in range(0, m, 2000):                                  # --variant-chunk default
        end = min(start + 2000, m)
        codes = ((payload[start:end, :, None] >> shifts) & 3).reshape(end - start, -1)[:, :n]
        X = np.ascontiguousarray(lut[codes].T)                       # [samples x variants] float32, C order
        ns = {"np": np, "X": X, "args": args, "hwe_pval": refpca.hwe_pval}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            exec(code, ns)
        keep[start:end] = ns["ok"]
        for k in parts:
            parts[k][start:end] = ns[k]
    return keep, parts


if __name__ == "__main__":
    main()
