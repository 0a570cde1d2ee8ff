"""CPU oracle for the genomic_pca hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain numpy (f64 / f32) restatement of the reference's
algorithm for the path named in BASELINE.json's north_star.  It is *the
checker*, never the product:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import it;
  * nothing under ``genomic_pca_b200/`` imports it, and the product path fails
    loudly when the CUDA library is missing.

PARITY UNPINNED.  The reference (`/root/reference`, Rust) cannot be compiled in
this image (no cargo/rustc), holds no unit tests, golden vectors or fixtures
with expected outputs, and the randomized-SVD / EigenSNP arithmetic lives in
the external crate ``efficient_pca`` (``Cargo.toml:30``, git branch ``main``,
unpinned, source not on disk).  What *is* pinned:

  * the in-tree integer/f64 statistics (decode, allele counts, QC ladder, HWE,
    mean/sigma, LD mapping, f32 standardisation) follow ``src/prepare.rs`` and
    ``src/vcf.rs`` line by line (each function cites file:line), and are
    cross-checked against the reference's own *Python* helpers that can be
    imported here (``tests/disk.py`` decode, ``tests/pca.py`` HWE) through
    ``tests/golden/make_golden.py``;
  * the PCA stages are checked against an exact f64 ``eigh`` PCA on structured
    synthetic data (the only ground truth available for the external crate).
"""
