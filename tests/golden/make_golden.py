#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's own data and importable Python helpers.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py

What comes from the REFERENCE itself (not from oracle/):
  * ``decode_ref``  -- tests/disk.py:89-137 ``process_snp_block_for_accessor_test`` run on
    rows of data/chr22_subset50.bed (its convention: 00->0, 10->1, 11->2, 01->255).
  * ``hwe_ref``     -- tests/pca.py:54-66 ``hwe_pval`` (an independent HWE chi-square in the
    reference repository; ``bed_reader`` is stubbed because only that function is used).
What is stored beside them: the packed rows they were computed from, so the tests can feed
the same bytes to oracle/ and to the CUDA path anywhere.
"""
import importlib.util
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


if __name__ == "__main__":
    main()
