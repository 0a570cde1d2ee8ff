"""ctypes binding of include/gpca.h (host-side mirror used by tests and bench.py).

Names follow the reference's host API where one exists:
  Context.load_bed / snp_counts / snp_qc          <-> MicroarrayDataPreparer (src/prepare.rs:922-1096, 1100-1422)
  Context.get_standardized_snp_sample_block       <-> PcaReadyGenotypeAccessor (src/prepare.rs:1838-2030)
  Context.rfit                                     <-> pca_runner::run_genomic_pca (src/main.rs:598-679)
  Context.eigensnp                                 <-> EigenSNPCoreAlgorithm::compute_pca (src/main.rs:365)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def library_path() -> str:
    return os.path.join(_HERE, "libgpca.so")


class GpcaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gpca error {code}: {msg}")
        self.code = code


def _load():
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    return C.CDLL(path)


lib = _load()

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p)


class QcConfig(C.Structure):
    _fields_ = [("min_call_rate", C.c_double), ("min_maf", C.c_double), ("max_hwe_p", C.c_double)]

    def __init__(self, min_call_rate=0.98, min_maf=0.01, max_hwe_p=1e-6):
        super().__init__(min_call_rate, min_maf, max_hwe_p)


class EigenSnpConfig(C.Structure):
    _fields_ = [("target_num_global_pcs", C.c_uint32), ("components_per_ld_block", C.c_uint32),
                ("subset_factor", C.c_double), ("min_subset_size", C.c_uint64), ("max_subset_size", C.c_uint64),
                ("global_oversampling", C.c_uint32), ("global_power_iters", C.c_uint32),
                ("local_oversampling", C.c_uint32), ("local_power_iters", C.c_uint32),
                ("random_seed", C.c_uint64), ("snp_processing_strip_size", C.c_uint32),
                ("refine_pass_count", C.c_uint32), ("collect_diagnostics", C.c_uint32)]

    def __init__(self, **kw):
        super().__init__()
        lib.gpca_eigensnp_default_cfg(C.byref(self))
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError(k)
            setattr(self, k, v)


def _sig(name, restype, *argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


_sig("gpca_init", C.c_int, C.POINTER(C.c_void_p), C.c_int)
_sig("gpca_destroy", None, C.c_void_p)
_sig("gpca_last_error", C.c_char_p, C.c_void_p)
_sig("gpca_version", C.c_char_p)
_sig("gpca_launch_count", C.c_uint64, C.c_void_p)
_sig("gpca_reset_launch_count", None, C.c_void_p)
_sig("gpca_set_sketch_engine", C.c_int, C.c_void_p, C.c_int)
_sig("gpca_set_batch_blocks", C.c_int, C.c_void_p, C.c_int)
_sig("gpca_last_sketch_engine", C.c_int, C.c_void_p)
_sig("gpca_ingest_bed_file", C.c_int, C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p,
     C.c_double, _u8p, _f32p, _f32p, _u8p, _u64p)
_sig("gpca_ingest_bed_file_rows", C.c_int, C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p,
     C.c_uint64, C.c_void_p, C.c_double, _u8p, _f32p, _f32p, _u8p, _u64p)
_sig("gpca_synth_bed_device", C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32,
     C.c_double, C.c_double, C.c_double)
_sig("gpca_sketch_stats", C.c_int, C.c_void_p, _f64p, _f64p, _u64p, C.c_int)
_sig("gpca_sketch_kernel_ms", C.c_double, C.c_void_p)
_sig("gpca_set_allreduce", C.c_int, C.c_void_p, ALLREDUCE_FN, C.c_void_p)
_sig("gpca_set_host_threads", C.c_int, C.c_void_p, C.c_uint32)
_sig("gpca_set_sketch_timing", C.c_int, C.c_void_p, C.c_int)
_sig("gpca_bind_host_to_device", C.c_int, C.c_void_p)
_sig("gpca_set_memory_reserve", C.c_int, C.c_void_p, C.c_uint64)
_sig("gpca_resident_snp_rows", C.c_uint64, C.c_void_p)
_sig("gpca_eigensnp_workspace_bytes", C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p)
_sig("gpca_set_ingest_mask", C.c_int, C.c_void_p, _u8p, C.c_uint64)
_sig("gpca_synth_bed_host", C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32,
     C.c_double, C.c_double, C.c_double)
_sig("gpca_count_kernel_ms", C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, _f64p)
_sig("gpca_host_alloc", C.c_void_p, C.c_void_p, C.c_uint64)
_sig("gpca_host_free", None, C.c_void_p, C.c_void_p, C.c_uint64)
_sig("gpca_comm_unique_id", C.c_int, _u8p)
_sig("gpca_comm_init", C.c_int, C.c_void_p, _u8p, C.c_int, C.c_int)
_sig("gpca_comm_finalize", C.c_int, C.c_void_p)
_sig("gpca_comm_world", C.c_int, C.c_void_p)
_sig("gpca_collective_count", C.c_uint64, C.c_void_p)
_sig("gpca_set_shard", C.c_int, C.c_void_p, C.c_uint64, C.c_uint64)
_sig("gpca_load_bed", C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, _i64p, C.c_uint64)
_sig("gpca_load_bed_device", C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64)
_sig("gpca_load_u8_variant_major", C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64)
_sig("gpca_num_samples", C.c_uint64, C.c_void_p)
_sig("gpca_num_snps", C.c_uint64, C.c_void_p)
_sig("gpca_num_pca_snps", C.c_uint64, C.c_void_p)
_sig("gpca_snp_counts", C.c_int, C.c_void_p, _u32p, _u32p, _u32p, _u32p)
_sig("gpca_snp_qc", C.c_int, C.c_void_p, C.POINTER(QcConfig), _u8p, _f32p, _f32p, _u8p)
_sig("gpca_vcf_maf_filter", C.c_int, C.c_void_p, C.c_double, _u8p, _f32p, _f32p)
_sig("gpca_hwe_chi_squared_p_value", C.c_double, C.c_uint64, C.c_uint64, C.c_uint64)
_sig("gpca_set_pca_snps", C.c_int, C.c_void_p, _u64p, C.c_uint64, _f32p, _f32p)
_sig("gpca_set_pca_snps_mask", C.c_int, C.c_void_p, _u8p, _f32p, _f32p, _u64p)
_sig("gpca_ingest_bed", C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p,
     C.c_double, _u8p, _f32p, _f32p, _u8p, _u64p)
_sig("gpca_get_standardized_block", C.c_int, C.c_void_p, _u64p, C.c_uint64, _u64p, C.c_uint64, _f32p)
_sig("gpca_sketch_snp_side", C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32)
_sig("gpca_sketch_sample_side", C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32)
_sig("gpca_dense_product", C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32,
     C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32)
_sig("gpca_synchronize", C.c_int, C.c_void_p)
_sig("gpca_get_stream", C.c_void_p, C.c_void_p)
_sig("gpca_rfit", C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, _f64p, _f64p, _f32p,
     _u32p)
_sig("gpca_eigensnp_default_cfg", None, C.POINTER(EigenSnpConfig))
_sig("gpca_eigensnp", C.c_int, C.c_void_p, C.POINTER(EigenSnpConfig), _u64p, C.c_uint64, _u64p, _f32p, _f64p, _f32p,
     _u32p)
_sig("gpca_eigensnp_diagnostics", C.c_char_p, C.c_void_p)
_sig("gpca_map_snps_to_ld_blocks", C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.c_uint64,
     C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_uint64, _i64p, _i64p, _u64p, _u64p, _u64p)


COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (one rank calls it, the host hands the bytes to the others)."""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    rc = lib.gpca_comm_unique_id(buf)
    if rc != 0:
        raise GpcaError(rc, "gpca_comm_unique_id: NCCL not available")
    return bytes(buf)


def eigensnp_workspace_bytes(n_samples: int, n_pca_snps: int, n_blocks: int, cfg=None) -> int:
    return int(lib.gpca_eigensnp_workspace_bytes(int(n_samples), int(n_pca_snps), int(n_blocks),
                                                 None if cfg is None else C.addressof(cfg)))


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def hwe_chi_squared_p_value(hom1, het, hom2) -> float:
    return float(lib.gpca_hwe_chi_squared_p_value(int(hom1), int(het), int(hom2)))


def map_snps_to_ld_blocks(snp_chrom, snp_bp, blk_chrom, blk_start, blk_end):
    """Host-side twin of MicroarrayDataPreparer::map_snps_to_ld_blocks (src/prepare.rs:1424-1563).
    Chromosome strings must already be normalised.  Returns (pca_pos, block_of, n_pca, n_blocks, order)."""
    nq, nb = len(snp_chrom), len(blk_chrom)
    sc = (C.c_char_p * max(nq, 1))(*[s.encode() for s in snp_chrom])
    bc = (C.c_char_p * max(nb, 1))(*[s.encode() for s in blk_chrom])
    sbp = np.ascontiguousarray(snp_bp, dtype=np.int32)
    bs = np.ascontiguousarray(blk_start, dtype=np.int32)
    be = np.ascontiguousarray(blk_end, dtype=np.int32)
    pca_pos = np.empty(max(nq, 1), dtype=np.int64)
    block_of = np.empty(max(nq, 1), dtype=np.int64)
    order = np.zeros(max(nb, 1), dtype=np.uint64)
    n_pca, n_blk = C.c_uint64(0), C.c_uint64(0)
    rc = lib.gpca_map_snps_to_ld_blocks(sc, _ptr(sbp, C.POINTER(C.c_int32)), nq, bc, _ptr(bs, C.POINTER(C.c_int32)),
                                        _ptr(be, C.POINTER(C.c_int32)), nb, _ptr(pca_pos, _i64p),
                                        _ptr(block_of, _i64p), C.byref(n_pca), C.byref(n_blk), _ptr(order, _u64p))
    if rc != 0:
        raise GpcaError(rc, "gpca_map_snps_to_ld_blocks")
    return pca_pos[:nq], block_of[:nq], int(n_pca.value), int(n_blk.value), order[:int(n_blk.value)]


class Context:
    """One GPU context (opaque gpca_ctx)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        rc = lib.gpca_init(C.byref(self._h), device)
        if rc != 0:
            raise GpcaError(rc, "gpca_init failed: no sm_100 (B200) device visible -- there is no CPU fallback")
        self._cb = None

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib.gpca_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc):
        if rc != 0:
            raise GpcaError(rc, (lib.gpca_last_error(self._h) or b"").decode())

    # -- configuration
    def set_sketch_engine(self, engine: int):
        self._chk(lib.gpca_set_sketch_engine(self._h, engine))

    def synth_bed_device(self, dev_ptr: int, n_samples: int, n_snps: int, snp_offset: int = 0, seed: int = 20260101,
                         n_pops: int = 22, fst: float = 0.1, missing_rate: float = 0.0, fst_grade: float = 0.0):
        """Benchmark input: structured synthetic genotypes written on the device in .bed layout (see gpca.h)."""
        self._chk(lib.gpca_synth_bed_device(self._h, dev_ptr, n_samples, n_snps, snp_offset, seed, n_pops, fst,
                                            missing_rate, fst_grade))

    def synth_bed_host(self, host_ptr: int, n_samples: int, n_snps: int, snp_offset: int = 0, seed: int = 20260101,
                       n_pops: int = 22, fst: float = 0.1, missing_rate: float = 0.0, fst_grade: float = 0.0):
        """The same rows as synth_bed_device, written to host memory (stands in for a .bed file read by the host)."""
        self._chk(lib.gpca_synth_bed_host(self._h, host_ptr, n_samples, n_snps, snp_offset, seed, n_pops, fst,
                                          missing_rate, fst_grade))

    def count_kernel_ms(self, dev_ptr: int, n_samples: int, n_snps: int, reps: int = 10) -> float:
        """mean device time (ms) of one launch of the allele-count kernel (K-a) on a device-resident payload"""
        ms = C.c_double(0.0)
        self._chk(lib.gpca_count_kernel_ms(self._h, dev_ptr, n_samples, n_snps, reps, C.byref(ms)))
        return float(ms.value)

    def set_host_threads(self, n: int):
        self._chk(lib.gpca_set_host_threads(self._h, int(n)))

    def bind_host_to_device(self) -> int:
        """Bind this thread (and the threads created from it) to the CPUs next to the GPU; returns their number."""
        return int(lib.gpca_bind_host_to_device(self._h))

    def set_sketch_timing(self, on: bool):
        self._chk(lib.gpca_set_sketch_timing(self._h, 1 if on else 0))

    def set_memory_reserve(self, nbytes: int):
        self._chk(lib.gpca_set_memory_reserve(self._h, int(nbytes)))

    @property
    def resident_snp_rows(self) -> int:
        """Rows of the SNP-major matrix that are resident (== num_pca_snps unless the two orientations did not fit)."""
        return int(lib.gpca_resident_snp_rows(self._h))

    def host_alloc(self, nbytes: int) -> int:
        """Pinned host buffer (huge pages, registered with CUDA); returns its address.  Free with host_free."""
        p = lib.gpca_host_alloc(self._h, int(nbytes))
        if not p:
            raise GpcaError(-5, (lib.gpca_last_error(self._h) or b"").decode())
        return int(p)

    def host_free(self, ptr: int, nbytes: int):
        lib.gpca_host_free(self._h, C.c_void_p(ptr), int(nbytes))

    def set_ingest_mask(self, mask):
        """Pre-selection of loaded rows for the next ingest (None clears it)."""
        if mask is None:
            self._chk(lib.gpca_set_ingest_mask(self._h, None, 0))
            return
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        self._chk(lib.gpca_set_ingest_mask(self._h, _ptr(mask, _u8p), mask.size))

    # -- the library's own exchange (NCCL)
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(unique_id)
        self._chk(lib.gpca_comm_init(self._h, buf, int(rank), int(world)))

    def comm_finalize(self):
        self._chk(lib.gpca_comm_finalize(self._h))

    @property
    def comm_world(self) -> int:
        return int(lib.gpca_comm_world(self._h))

    @property
    def eigensnp_diagnostics(self) -> str:
        """JSON record of the last eigensnp() call made with collect_diagnostics=1 ("" if none)."""
        return (lib.gpca_eigensnp_diagnostics(self._h) or b"").decode()

    @property
    def collective_count(self) -> int:
        return int(lib.gpca_collective_count(self._h))

    @property
    def last_sketch_engine(self) -> int:
        """the engine the last sketch pass actually ran on (0 SIMT, 1 tcgen05 f16, 2 tcgen05 i8)"""
        return int(lib.gpca_last_sketch_engine(self._h))

    def set_batch_blocks(self, on: bool):
        self._chk(lib.gpca_set_batch_blocks(self._h, 1 if on else 0))

    def set_shard(self, offset: int, total: int):
        self._chk(lib.gpca_set_shard(self._h, offset, total))

    def set_allreduce(self, fn):
        """fn(dev_ptr:int, count:int, dtype:int(0=f32,1=f64), stream:int) -> None"""
        if fn is None:
            self._cb = None
            self._chk(lib.gpca_set_allreduce(self._h, C.cast(None, ALLREDUCE_FN), None))
            return

        def tramp(ptr, count, dtype, stream, user):
            try:
                fn(ptr, count, dtype, stream)
                return 0
            except Exception as e:  # pragma: no cover
                print("allreduce hook raised:", e)
                return 1

        self._cb = ALLREDUCE_FN(tramp)
        self._chk(lib.gpca_set_allreduce(self._h, self._cb, None))

    @property
    def launch_count(self) -> int:
        return int(lib.gpca_launch_count(self._h))

    def reset_launch_count(self):
        lib.gpca_reset_launch_count(self._h)

    def sketch_stats(self, reset=False):
        ms, by, n = C.c_double(), C.c_double(), C.c_uint64()
        self._chk(lib.gpca_sketch_stats(self._h, C.byref(ms), C.byref(by), C.byref(n), int(reset)))
        self.last_kernel_ms = float(lib.gpca_sketch_kernel_ms(self._h))
        return ms.value, by.value, int(n.value)

    # -- ingest
    def load_bed(self, payload: np.ndarray, n_samples: int, n_snps: int, keep_samples=None):
        payload = np.ascontiguousarray(payload, dtype=np.uint8)
        assert payload.size == ((n_samples + 3) // 4) * n_snps, "payload size != n_snps*ceil(n_samples/4)"
        keep = None if keep_samples is None else np.ascontiguousarray(keep_samples, dtype=np.int64)
        self._chk(lib.gpca_load_bed(self._h, payload.ctypes.data, n_samples, n_snps, _ptr(keep, _i64p),
                                    0 if keep is None else keep.size))

    def load_bed_device(self, dev_ptr: int, n_samples: int, n_snps: int):
        self._chk(lib.gpca_load_bed_device(self._h, dev_ptr, n_samples, n_snps))

    def load_bed_host_ptr(self, host_ptr: int, n_samples: int, n_snps: int):
        self._chk(lib.gpca_load_bed(self._h, host_ptr, n_samples, n_snps, None, 0))

    def load_u8_variant_major(self, dosage: np.ndarray):
        dosage = np.ascontiguousarray(dosage, dtype=np.uint8)
        d, n = dosage.shape
        self._chk(lib.gpca_load_u8_variant_major(self._h, dosage.ctypes.data, n, d))

    @property
    def num_samples(self):
        return int(lib.gpca_num_samples(self._h))

    @property
    def num_snps(self):
        return int(lib.gpca_num_snps(self._h))

    @property
    def num_pca_snps(self):
        return int(lib.gpca_num_pca_snps(self._h))

    # -- statistics
    def snp_counts(self):
        m = self.num_snps
        out = [np.empty(max(m, 1), dtype=np.uint32) for _ in range(4)]
        self._chk(lib.gpca_snp_counts(self._h, *[_ptr(a, _u32p) for a in out]))
        return tuple(a[:m] for a in out)

    def ingest_bed(self, payload, n_samples: int, n_snps: int, qc: QcConfig | None = None, vcf_maf: float = 0.01,
                   keep_samples=None, want_stats=True, out=None):
        """Load + filter + build the resident matrices in one streaming pass (gpca_ingest_bed).
        `payload`: numpy uint8 array (SNP-major .bed payload without the magic) or an integer host address (e.g. a
        pinned buffer).  `qc` given -> the PLINK QC ladder; otherwise the VCF MAF filter at `vcf_maf`.
        `out`: optional preallocated (keep u8, mean f32, sd f32[, fail_code u8]) arrays of n_snps elements to fill
        (returned as they are, keep as uint8) -- saves fresh allocations when the call is repeated.
        Returns (keep, mean, sd, fail_code | None, n_pca)."""
        if isinstance(payload, (int, np.integer)):
            ptr = int(payload)
        else:
            payload = np.ascontiguousarray(payload, dtype=np.uint8)
            ptr = payload.ctypes.data
        ks = None if keep_samples is None else np.ascontiguousarray(keep_samples, dtype=np.int64)
        m = int(n_snps)
        n = C.c_uint64(0)
        if out is not None:
            keep, mean, sd = out[0], out[1], out[2]
            code = out[3] if len(out) > 3 else None
            assert keep.dtype == np.uint8 and mean.dtype == np.float32 and sd.dtype == np.float32
            assert keep.size >= m and mean.size >= m and sd.size >= m and (code is None or code.size >= m)
            self._chk(lib.gpca_ingest_bed(self._h, ptr, n_samples, n_snps, None if ks is None else ks.ctypes.data,
                                          0 if ks is None else ks.size, None if qc is None else C.addressof(qc),
                                          float(vcf_maf), _ptr(keep, _u8p), _ptr(mean, _f32p), _ptr(sd, _f32p),
                                          _ptr(code, _u8p), C.byref(n)))
            return keep, mean, sd, code, int(n.value)
        keep = np.empty(max(m, 1), dtype=np.uint8) if want_stats else None
        mean = np.empty(max(m, 1), dtype=np.float32) if want_stats else None
        sd = np.empty(max(m, 1), dtype=np.float32) if want_stats else None
        code = np.empty(max(m, 1), dtype=np.uint8) if (want_stats and qc is not None) else None
        self._chk(lib.gpca_ingest_bed(self._h, ptr, n_samples, n_snps, None if ks is None else ks.ctypes.data,
                                      0 if ks is None else ks.size, None if qc is None else C.addressof(qc),
                                      float(vcf_maf), _ptr(keep, _u8p), _ptr(mean, _f32p), _ptr(sd, _f32p),
                                      _ptr(code, _u8p), C.byref(n)))
        if not want_stats:
            return None, None, None, None, int(n.value)
        return keep[:m].astype(bool), mean[:m], sd[:m], (None if code is None else code[:m]), int(n.value)

    def ingest_bed_file(self, bed_path: str, n_samples: int, n_snps: int, qc: QcConfig | None = None,
                        vcf_maf: float = 0.01, keep_samples=None):
        """gpca_ingest_bed_file: the streaming ingest straight from a .bed file (magic and size checked)."""
        ks = None if keep_samples is None else np.ascontiguousarray(keep_samples, dtype=np.int64)
        m = int(n_snps)
        keep = np.empty(max(m, 1), dtype=np.uint8)
        mean = np.empty(max(m, 1), dtype=np.float32)
        sd = np.empty(max(m, 1), dtype=np.float32)
        code = np.empty(max(m, 1), dtype=np.uint8) if qc is not None else None
        n = C.c_uint64(0)
        self._chk(lib.gpca_ingest_bed_file(self._h, os.fsencode(bed_path), n_samples, n_snps,
                                           None if ks is None else ks.ctypes.data, 0 if ks is None else ks.size,
                                           None if qc is None else C.addressof(qc), float(vcf_maf), _ptr(keep, _u8p),
                                           _ptr(mean, _f32p), _ptr(sd, _f32p), _ptr(code, _u8p), C.byref(n)))
        return keep[:m].astype(bool), mean[:m], sd[:m], (None if code is None else code[:m]), int(n.value)

    def snp_qc(self, cfg: QcConfig | None = None):
        cfg = cfg or QcConfig()
        m = self.num_snps
        keep = np.empty(max(m, 1), dtype=np.uint8)
        mean = np.empty(max(m, 1), dtype=np.float32)
        sd = np.empty(max(m, 1), dtype=np.float32)
        code = np.empty(max(m, 1), dtype=np.uint8)
        self._chk(lib.gpca_snp_qc(self._h, C.byref(cfg), _ptr(keep, _u8p), _ptr(mean, _f32p), _ptr(sd, _f32p),
                                  _ptr(code, _u8p)))
        return keep[:m].astype(bool), mean[:m], sd[:m], code[:m]

    def vcf_maf_filter(self, maf=0.01):
        m = self.num_snps
        keep = np.empty(max(m, 1), dtype=np.uint8)
        mean = np.empty(max(m, 1), dtype=np.float32)
        sd = np.empty(max(m, 1), dtype=np.float32)
        self._chk(lib.gpca_vcf_maf_filter(self._h, float(maf), _ptr(keep, _u8p), _ptr(mean, _f32p), _ptr(sd, _f32p)))
        return keep[:m].astype(bool), mean[:m], sd[:m]

    def set_pca_snps(self, snp_idx, mean, sd):
        idx = np.ascontiguousarray(snp_idx, dtype=np.uint64)
        mean = np.ascontiguousarray(mean, dtype=np.float32)
        sd = np.ascontiguousarray(sd, dtype=np.float32)
        assert idx.size == mean.size == sd.size
        self._chk(lib.gpca_set_pca_snps(self._h, _ptr(idx, _u64p), idx.size, _ptr(mean, _f32p), _ptr(sd, _f32p)))

    def set_pca_snps_mask(self, keep, mean_all, sd_all) -> int:
        keep = np.ascontiguousarray(keep, dtype=np.uint8)
        mean_all = np.ascontiguousarray(mean_all, dtype=np.float32)
        sd_all = np.ascontiguousarray(sd_all, dtype=np.float32)
        assert keep.size == mean_all.size == sd_all.size == self.num_snps
        n = C.c_uint64(0)
        self._chk(lib.gpca_set_pca_snps_mask(self._h, _ptr(keep, _u8p), _ptr(mean_all, _f32p), _ptr(sd_all, _f32p),
                                             C.byref(n)))
        return int(n.value)

    def get_standardized_snp_sample_block(self, pca_snp_ids, qc_sample_ids=None):
        ids = np.ascontiguousarray(pca_snp_ids, dtype=np.uint64)
        samp = None if qc_sample_ids is None else np.ascontiguousarray(qc_sample_ids, dtype=np.uint64)
        ns = self.num_samples if samp is None else samp.size
        out = np.zeros((ids.size, ns), dtype=np.float32)
        self._chk(lib.gpca_get_standardized_block(self._h, _ptr(ids, _u64p), ids.size, _ptr(samp, _u64p), ns,
                                                  _ptr(out, _f32p)))
        return out

    # -- sketch passes on device pointers
    def sketch_snp_side(self, in_ptr: int, out_ptr: int, l: int, ld: int):
        self._chk(lib.gpca_sketch_snp_side(self._h, in_ptr, out_ptr, l, ld))

    def sketch_sample_side(self, in_ptr: int, out_ptr: int, l: int, ld: int):
        self._chk(lib.gpca_sketch_sample_side(self._h, in_ptr, out_ptr, l, ld))

    def dense_product(self, c_ptr, n, r, ldc, cols_mode, in_ptr, l, ld, out_ptr, ldo, f=0, e=0, a=0, b=0):
        """gpca_dense_product on device pointers (0 = NULL for the optional scale vectors)."""
        self._chk(lib.gpca_dense_product(self._h, c_ptr, n, r, ldc, 1 if cols_mode else 0, in_ptr, l, ld, f or None,
                                         e or None, a or None, b or None, out_ptr, ldo))

    @property
    def stream(self) -> int:
        return int(lib.gpca_get_stream(self._h) or 0)

    def synchronize(self):
        self._chk(lib.gpca_synchronize(self._h))

    # -- drivers
    def rfit(self, k, oversample=10, power_iters=2, seed=None, want_loadings=True, out=None, want_scores=True):
        """`out` = (scores f64 [N, k], eigenvalues f64 [k], loadings f32 [D, k] or None): caller-owned result buffers,
        as the C ABI has them (a host that calls repeatedly allocates them once; fresh `np.zeros` arrays are mapped
        lazily and page-fault while the results land -- 20,000 faults for the 500,000 x 20 f64 scores).
        `want_scores=False` passes NULL for the scores (a shard of a multi-GPU run whose copy of the -- identical --
        scores nobody reads: the 40 MB download and the widening to f64 are then skipped)."""
        n, d = self.num_samples, self.num_pca_snps
        kk = max(1, min(k, n))
        if out is not None:
            scores, ev, load = out
            assert scores is None or (scores.dtype == np.float64 and scores.shape == (n, kk) and scores.flags.c_contiguous)
            assert ev.dtype == np.float64 and ev.shape == (kk,)
            assert load is None or (load.dtype == np.float32 and load.shape == (d, kk) and load.flags.c_contiguous)
        else:
            scores = np.zeros((n, kk), dtype=np.float64) if want_scores else None
            ev = np.zeros(kk, dtype=np.float64)
            load = np.zeros((d, kk), dtype=np.float32) if want_loadings else None
        if not want_scores:
            scores = None
        kout = C.c_uint32(0)
        self._chk(lib.gpca_rfit(self._h, k, oversample, power_iters, 0 if seed is None else int(seed),
                                0 if seed is None else 1, _ptr(scores, _f64p), _ptr(ev, _f64p), _ptr(load, _f32p),
                                C.byref(kout)))
        ko = int(kout.value)
        if ko != kk:
            if scores is not None:
                scores = scores.reshape(-1)[:n * ko].reshape(n, ko)
            ev = ev[:ko]
            if load is not None:
                load = load.reshape(-1)[:d * ko].reshape(d, ko)
        return scores, ev, load

    def eigensnp(self, block_snp_ids, cfg: EigenSnpConfig | None = None, out=None, want_scores=True):
        """`out` = (scores f32 [N, k] or None, eigenvalues f64 [k], loadings f32 [D, k]): caller-owned result buffers
        (see rfit; `want_scores=False` as there)."""
        cfg = cfg or EigenSnpConfig()
        n, d = self.num_samples, self.num_pca_snps
        offs = np.zeros(len(block_snp_ids) + 1, dtype=np.uint64)
        for i, b in enumerate(block_snp_ids):
            offs[i + 1] = offs[i] + len(b)
        flat = (np.concatenate([np.asarray(b, dtype=np.uint64) for b in block_snp_ids])
                if len(block_snp_ids) else np.zeros(0, dtype=np.uint64))
        flat = np.ascontiguousarray(flat, dtype=np.uint64)
        k = int(cfg.target_num_global_pcs)
        if out is not None:
            scores, ev, load = out
            assert scores is None or (scores.dtype == np.float32 and scores.shape == (n, k) and scores.flags.c_contiguous)
            assert ev.dtype == np.float64 and ev.shape == (k,)
            assert load.dtype == np.float32 and load.shape == (d, k) and load.flags.c_contiguous
        else:
            scores = np.zeros((n, k), dtype=np.float32) if want_scores else None
            ev = np.zeros(k, dtype=np.float64)
            load = np.zeros((d, k), dtype=np.float32)
        if not want_scores:
            scores = None
        kout = C.c_uint32(0)
        self._chk(lib.gpca_eigensnp(self._h, C.byref(cfg), _ptr(offs, _u64p), len(block_snp_ids), _ptr(flat, _u64p),
                                    _ptr(scores, _f32p), _ptr(ev, _f64p), _ptr(load, _f32p), C.byref(kout)))
        ko = int(kout.value)
        if ko != k:
            if scores is not None:
                scores = scores.reshape(-1)[:n * ko].reshape(n, ko)
            ev = ev[:ko]
            load = load.reshape(-1)[:d * ko].reshape(d, ko)
        return scores, ev, load
