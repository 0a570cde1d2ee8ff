#!/usr/bin/env python
"""bench.py -- the hot path (randomized-SVD sketch passes of the rfit / EigenSNP PCA) on synthetic data.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload = BASELINE.json config 4, the configuration the metric is quoted on: 500,000 samples x 700,000 SNPs, 2-bit
packed (87.5 GB), k = 20, STRONG-scaled over the GPUs -- rank r owns the SNPs of LD blocks [r B/N, (r+1) B/N) of
B = 1,696 blocks; the one exchange is the sum of the N x l sketch after every sample-side pass (NCCL, issued by the
library on its own stream).
One "step" = one complete rfit PCA (k = 20, oversample 10, q = 2: 7 sketch passes over the resident 2-bit matrix, the
re-orthonormalisations, the small eigensolves, scores back on the host).
metric = genotype GB/s per sketch pass = packed bytes all ranks stream per pass x passes / step time.
  value : matrix resident in HBM when the timed region starts (CUDA events on the library's stream, max over ranks)
  e2e   : the same PCA through the C ABI from HOST buffers: gpca_ingest_bed of each rank's shard from pinned host memory
          (H2D, counts, QC ladder on host threads, both resident orientations) + gpca_rfit + scores on the host.
Extra records in the same line: "eigensnp" (the EigenSNP mode of the same configuration, device-resident and e2e) and
"c3" (BASELINE config 3, 2,504 x 10M rfit on rank 0's GPU: the round-1 headline, with `parity_at_scale` against an
exact f64 eigen-decomposition of the GRM).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DATA_SEED = 20260101
K_COMPONENTS = 20
OVERSAMPLE = 10
POWER_ITERS = 2
RFIT_SEED = 42
N_POPS = K_COMPONENTS + 2
FST, FST_GRADE = 0.1, 1.0        # graded drift: distinct structural eigenvalues (gpca.h, gpca_synth_bed_device)
C4_SAMPLES, C4_SNPS, C4_BLOCKS = 500_000, 700_000, 1696
C3_SAMPLES, C3_SNPS = 2504, 10_000_000


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=None, help="timed rfit steps (default 10; 2 for --impl reference)")
    p.add_argument("--warmup", type=int, default=None, help="untimed steps (default 3; 0 for --impl reference)")
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--samples", type=int, default=C4_SAMPLES)
    p.add_argument("--snps", type=int, default=C4_SNPS, help="SNPs of the WHOLE job (strong scaling: split over the GPUs)")
    p.add_argument("--blocks", type=int, default=C4_BLOCKS, help="LD blocks of the whole job")
    p.add_argument("--engine", type=int, default=None)
    p.add_argument("--components", type=int, default=None, help="k (default 20)")
    p.add_argument("--power-iters", type=int, default=None, help="q (default 2)")
    p.add_argument("--cpu-snps", type=int, default=None,
                   help="SNP rows of the CPU sample (default 512; --impl reference: up to 1024, scaled down so that\n"
                        "steps + warmup stay within ~2 minutes of CPU time)")
    p.add_argument("--e2e-steps", type=int, default=2)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-eigensnp", action="store_true")
    p.add_argument("--no-c3", action="store_true", help="skip the config-3 record")
    p.add_argument("--no-parity", action="store_true", help="skip parity_at_scale of the config-3 record")
    p.add_argument("--c3-snps", type=int, default=C3_SNPS)
    a = p.parse_args()
    if a.steps is None:
        a.steps = 2 if a.impl == "reference" else 10
    if a.warmup is None:
        a.warmup = 0 if a.impl == "reference" else 3
    if a.cpu_snps is None:
        # one rfit over 500,000 samples x 1,024 SNPs takes ~20 s on 8-16 host cores (most of it LAPACK QR of N x l)
        a.cpu_snps = 512 if a.impl != "reference" else max(64, min(1024, int(1024 * 6 / max(a.steps + a.warmup, 1))))
    global K_COMPONENTS, POWER_ITERS, N_POPS
    if a.components is not None:
        K_COMPONENTS = a.components
        N_POPS = K_COMPONENTS + 2
    if a.power_iters is not None:
        POWER_ITERS = a.power_iters
    return a


# ------------------------------------------------------------------------------------------
def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region.  NVML is polled every 5 ms from a thread
    (one rfit step is ~15 ms: `nvidia-smi -lms 100` sees one or two samples of a default run); nvidia-smi is the
    fallback when pynvml is missing."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.nvml = None
        self.stop_flag = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES-relative index -> NVML handle through the PCI bus id
            import torch
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(self.index), "pci_bus_id") else None
            self.h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hh = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(hh).bus == bus:
                        self.h = hh
                        break
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml = pynvml
            self.thr = threading.Thread(target=self._poll, daemon=True)
            self.thr.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _poll(self):
        # (an NVML query takes milliseconds while the GPU is busy: two queries per sample, the power every fourth one,
        #  the maximum clock once)
        nv = self.nvml
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        it, pw = 0, None
        while not self.stop_flag:
            try:
                if it % 4 == 0:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), mx,
                                  nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h), pw))
            except Exception:
                pass
            it += 1
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thr.join(timeout=1.0)
            nv = self.nvml
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            reasons = sorted(nm for nm, bit in bits.items() if any(r[2] & bit for r in self.rows))
            sm = [r[0] for r in self.rows]
            pw = [r[3] for r in self.rows if r[3] is not None]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": float(min(sm)) if sm else None,
                    "sm_max_mhz": float(max(r[1] for r in self.rows if r[1] is not None)) if any(
                        r[1] is not None for r in self.rows) else None, "reasons": reasons,
                    "power_w": float(np.median(pw)) if pw else None, "samples": len(sm), "source": "nvml, 2 ms"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}



# ------------------------------------------------------------------------------------------
def synth_bed_device(torch, n_samples, n_snps, snp_offset, device, fst_grade=FST_GRADE):
    """Balding-Nichols genotypes generated on the device straight into PLINK .bed layout (P = k+2
    populations, ancestral AF ~ U(0.05,0.5), F_ST = 0.1 graded over the populations, no missing calls) by the library's
    counter-based generator (Philox keyed by (DATA_SEED, global SNP index, sample)): rows [a, b) of any shard are the
    rows [a, b) of the whole matrix, whatever the number of GPUs.  (tests / tools; the bench itself draws the payload
    into host memory, HostPayload.)"""
    import genomic_pca_b200 as gp
    bps = (n_samples + 3) // 4
    out = torch.empty((n_snps, bps), dtype=torch.uint8, device=device)
    gen = gp.Context(device.index or 0)
    gen.synth_bed_device(out.data_ptr(), n_samples, n_snps, snp_offset, DATA_SEED, N_POPS, FST, 0.0, fst_grade)
    gen.close()
    return out


def shard_of(rank, world, n_blocks_total, m_total):
    """Strong scaling: rank's LD blocks [b0, b1) of the whole job and the SNP rows [s0, s1) they cover."""
    edges = np.linspace(0, m_total, n_blocks_total + 1).astype(np.int64)
    b0, b1 = rank * n_blocks_total // world, (rank + 1) * n_blocks_total // world
    return b0, b1, int(edges[b0]), int(edges[b1])


def cpu_rfit_sample(n_samples, n_snps, steps=1, warmup=0):
    """The reference's CPU algorithm (oracle restatement: f64 matrix, OpenBLAS GEMMs, numpy QR/eigh)
    timed on a bounded sample of the same workload (the sample is drawn once, outside the timed region).
    Returns (seconds per step, passes, bytes per pass)."""
    from oracle import pca, synth
    g = np.concatenate([synth.balding_nichols(n_samples, min(512, n_snps - c0), n_pops=N_POPS, seed=7 + c0)[0]
                        for c0 in range(0, n_snps, 512)])
    mean = g.mean(1)
    sd = g.std(1, ddof=1)
    ok = sd > 1e-9
    g = g[ok]
    for _ in range(warmup):
        pca.rfit(pca.standardize_dense(g, mean[ok], sd[ok]), K_COMPONENTS, OVERSAMPLE, seed=RFIT_SEED, power_iters=POWER_ITERS)
    t0 = time.perf_counter()
    for _ in range(steps):
        # build_matrix (vcf.rs:317-345: u8 -> f64) + standardise + rfit + transform, as the reference does per run
        S = pca.standardize_dense(g, mean[ok], sd[ok])
        pca.rfit(S, K_COMPONENTS, OVERSAMPLE, seed=RFIT_SEED, power_iters=POWER_ITERS)
    dt = (time.perf_counter() - t0) / steps
    passes = 2 * POWER_ITERS + 3
    return dt, passes, g.shape[0] * ((n_samples + 3) // 4)


def blas_threads(n):
    """torchrun exports OMP_NUM_THREADS=1 for nproc > 1: the CPU arm sets its BLAS threads itself and records them."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=n)
    except Exception:
        return None


def run_reference(args, rank):
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lim = blas_threads(cores)
    n = args.samples
    dt, passes, bpp = cpu_rfit_sample(n, args.cpu_snps, steps=args.steps, warmup=min(args.warmup, 1))
    val = passes * bpp / dt / 1e9
    line = {
        "impl": "reference", "metric": "genotype_GBps_per_sketch_pass", "value": val, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"rfit k={K_COMPONENTS} oversample={OVERSAMPLE} q={POWER_ITERS}: {n} samples x "
                               f"{args.cpu_snps} SNPs (bounded CPU sample of BASELINE config 4, {n} x {args.snps})",
                   "passes_per_step": passes},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port",
                         "blas_threads": cores if lim is not None else os.environ.get("OMP_NUM_THREADS", "default"),
                         "sample": f"{n} samples x {args.cpu_snps} SNPs, numpy/OpenBLAS f64 restatement "
                                   "(reference binary cannot be built here: no cargo/rustc)"},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
class HostPayload:
    """The rank's rows of the .bed payload in pinned host memory (what a host would have read from the file)."""

    def __init__(self, ctx, n, m, snp_offset):
        self.ctx, self.n, self.m = ctx, n, m
        self.bps = (n + 3) // 4
        self.nbytes = self.m * self.bps
        t0 = time.perf_counter()
        self.ptr = ctx.host_alloc(self.nbytes)
        self.alloc_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        ctx.synth_bed_host(self.ptr, n, m, snp_offset, DATA_SEED, N_POPS, FST, 0.0, FST_GRADE)
        self.fill_s = time.perf_counter() - t0

    def free(self):
        if self.ptr:
            self.ctx.host_free(self.ptr, self.nbytes)
            self.ptr = 0


class PinnedArrays:
    """Caller-owned result buffers in pinned host memory (gpca_host_alloc): the library then lands scores / loadings
    straight off the bus instead of through its pinned staging buffers and host threads."""

    def __init__(self, ctx):
        self.ctx, self.held = ctx, []

    def ones(self, shape, dtype):
        nbytes = max(int(np.prod(shape)) * np.dtype(dtype).itemsize, 8)
        ptr = self.ctx.host_alloc(nbytes)
        self.held.append((ptr, nbytes))
        arr = np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(ptr))[:int(np.prod(shape)) * np.dtype(dtype).itemsize]
        arr = arr.view(dtype).reshape(shape)
        arr[...] = 1
        return arr

    def free(self):
        for ptr, nbytes in self.held:
            self.ctx.host_free(ptr, nbytes)
        self.held = []


def max_over_ranks(torch, dist, dev, world, vals):
    t = torch.tensor(vals, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def sum_over_ranks(torch, dist, dev, world, vals):
    t = torch.tensor(vals, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def timed_rfit(torch, dist, ctx, dev, world, steps, warmup, rfit_out, sample_clocks=None, want_scores=True):
    """K rfit calls on the resident matrix; device time between CUDA events on the library's stream, max over ranks."""
    def step():
        return ctx.rfit(K_COMPONENTS, OVERSAMPLE, power_iters=POWER_ITERS, seed=RFIT_SEED, want_loadings=False,
                        out=rfit_out, want_scores=want_scores)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    ctx.sketch_stats(reset=True)
    ctx.reset_launch_count()
    coll0 = ctx.collective_count
    barrier()
    if sample_clocks is not None:
        sample_clocks.start()
    lib_stream = torch.cuda.ExternalStream(ctx.stream, device=dev)      # the stream the kernels are launched on
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(lib_stream)
    t0 = time.perf_counter()
    for _ in range(steps):
        sc, ev, _ = step()
    ev1.record(lib_stream)
    barrier()
    wall = time.perf_counter() - t0
    dev_s = ev0.elapsed_time(ev1) * 1e-3
    clocks = sample_clocks.stop() if sample_clocks is not None else None
    launches = ctx.launch_count
    sk_ms, sk_bytes, sk_n = ctx.sketch_stats(reset=True)
    sk_kernel_ms = ctx.last_kernel_ms
    t_step = max_over_ranks(torch, dist, dev, world, [dev_s / steps])[0]
    return {"t_step": t_step, "wall_per_step": wall / steps, "clocks": clocks, "launches": launches, "sk_ms": sk_ms,
            "engine": ctx.last_sketch_engine,
            "sk_n": sk_n, "sk_kernel_ms": sk_kernel_ms, "eigenvalues": ev, "scores": sc,
            "collectives_per_step": (ctx.collective_count - coll0) / steps}


def e2e_loop(torch, dist, dev, world, fn, steps):
    """wall clock around `steps` calls of fn (each: ingest from pinned host memory + PCA + results on the host),
    barrier + synchronize on both sides, max over ranks"""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    barrier()
    te = (time.perf_counter() - t0) / steps
    return max_over_ranks(torch, dist, dev, world, [te])[0]


def sampled_parity(torch, ctx, payload, n, keep, mean, sd, dev, l=K_COMPONENTS + OVERSAMPLE, n_rows=192, n_cols=96, seed=7):
    """Parity of the hot kernel AT THE FULL SIZE of the run, by a size-independent route: both sketch orientations
    (gpca_sketch_snp_side: S B, gpca_sketch_sample_side: S^T W, seeded Gaussian operands) are run over the whole resident
    matrix, and a random sample of OUTPUT rows is recomputed in float64 by numpy from the genotypes decoded out of the
    host payload (`payload`: uint8 [M x ceil(N/4)], PLINK coding; count_a1 decode 00 -> 2, 01 -> missing, 10 -> 1,
    11 -> 0, src/prepare.rs:622-629) with the library's own mean / sd widened to f64 -- no kernel of the library on the
    checking side.  Returns the relative L2 errors; the integer engine's only rounding is the 16-bit fixed point of the
    dense operand (2^-15 of its maximum), so 1e-3 is a loose bound and ~1e-4 is what it measures."""
    rng = np.random.default_rng(seed)
    kept_idx = np.nonzero(np.asarray(keep))[0]
    d = kept_idx.size
    mu = np.asarray(mean)[kept_idx].astype(np.float64)
    isd = 1.0 / np.asarray(sd)[kept_idx].astype(np.float64)
    lut = np.array([2.0, np.nan, 1.0, 0.0])
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    b = torch.randn(n, l, device=dev, generator=g)
    w = torch.randn(d, l, device=dev, generator=g)
    out_s = torch.empty(d, l, device=dev)
    out_n = torch.empty(n, l, device=dev)
    torch.cuda.synchronize()
    ctx.sketch_snp_side(b.data_ptr(), out_s.data_ptr(), l, l)
    ctx.sketch_sample_side(w.data_ptr(), out_n.data_ptr(), l, l)
    ctx.synchronize()
    # snp side: sampled SNP rows, all samples
    rows = np.sort(rng.choice(d, size=min(n_rows, d), replace=False))
    by = payload[kept_idx[rows]]                                             # [rows x bps]
    codes = ((by[:, :, None] >> np.array([0, 2, 4, 6], dtype=np.uint8)) & 3).reshape(rows.size, -1)[:, :n]
    s_rows = np.nan_to_num((lut[codes] - mu[rows, None]) * isd[rows, None])   # a missing call standardises to 0
    ref = s_rows @ b.cpu().double().numpy()
    got = out_s[torch.as_tensor(rows, device=dev)].cpu().double().numpy()
    e_snp = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    # sample side: sampled samples (columns of the payload), all kept SNPs
    cols = np.sort(rng.choice(n, size=min(n_cols, n), replace=False))
    by = payload[np.ix_(kept_idx, cols // 4)]                                # [d x cols]
    codes = (by >> (2 * (cols % 4)).astype(np.uint8)[None, :]) & 3
    s_cols = np.nan_to_num((lut[codes] - mu[:, None]) * isd[:, None])
    ref = s_cols.T @ w.cpu().double().numpy()
    got = out_n[torch.as_tensor(cols, device=dev)].cpu().double().numpy()
    e_smp = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    del b, w, out_s, out_n
    return {"snp_side_rel_l2": e_snp, "sample_side_rel_l2": e_smp, "output_rows_checked": [int(rows.size), int(cols.size)],
            "l": int(l), "tolerance": 1e-3, "ok": bool(e_snp < 1e-3 and e_smp < 1e-3),
            "reference": "numpy float64 on genotypes decoded from the host payload (sampled output rows of both orientations)"}


def subspace_angle(a, b):
    """largest principal angle (rad) between the column spaces of two [n x k] matrices (float64 numpy)"""
    qa, _ = np.linalg.qr(np.asarray(a, dtype=np.float64))
    qb, _ = np.linalg.qr(np.asarray(b, dtype=np.float64))
    resid = qa - qb @ (qb.T @ qa)
    s = np.linalg.svd(resid, compute_uv=False)
    return float(np.arcsin(min(1.0, s[0])))


def exact_pca_f64(torch, payload, n, keep, mean, sd, k, chunk=32768):
    """Exact PCA of the standardized matrix for parity_at_scale: the N x N Gram matrix S^T S accumulated in float64 by
    torch over chunks of SNP rows decoded from the .bed payload (count_a1: 00 -> 2, 10 -> 1, 11 -> 0), then eigh.
    Test infrastructure, independent of the library's kernels.  Returns (explained variance [k], V [n x k])."""
    dev = payload.device
    bps = payload.shape[1]
    lut = torch.tensor([2.0, float("nan"), 1.0, 0.0], dtype=torch.float64, device=dev)
    shifts = torch.tensor([0, 2, 4, 6], dtype=torch.uint8, device=dev)
    idx = torch.as_tensor(np.nonzero(keep)[0], device=dev)
    mu = torch.as_tensor(mean[keep].astype(np.float64), device=dev)
    isd = 1.0 / torch.as_tensor(sd[keep].astype(np.float64), device=dev)
    gram = torch.zeros((n, n), dtype=torch.float64, device=dev)
    for c0 in range(0, idx.numel(), chunk):
        rows = idx[c0:c0 + chunk]
        b = payload.index_select(0, rows)
        codes = ((b.unsqueeze(-1) >> shifts) & 3).reshape(rows.numel(), bps * 4)[:, :n].long()
        s = (lut[codes] - mu[c0:c0 + chunk, None]) * isd[c0:c0 + chunk, None]
        gram.addmm_(s.t(), s)
        del b, codes, s
    evals, evecs = torch.linalg.eigh(gram)
    top = torch.argsort(evals, descending=True)[:k]
    return (evals[top] / (n - 1)).cpu().numpy(), evecs[:, top].cpu().numpy()


# ------------------------------------------------------------------------------------------
def run_c3(args, torch, gp, dev, pk, pk_src):
    """BASELINE config 3 (1000G shape, 2,504 x 10M, rfit k=20) on one GPU: device-resident step time, the sketch
    kernel's roofline, e2e from pinned host memory, and parity_at_scale against the exact f64 PCA."""
    n, m = C3_SAMPLES, args.c3_snps
    bps = (n + 3) // 4
    ctx = gp.Context(dev.index or 0)
    if args.engine is not None:
        ctx.set_sketch_engine(args.engine)
    ctx.set_sketch_timing(True)
    host = HostPayload(ctx, n, m, 0)
    pinned = PinnedArrays(ctx)
    stats_out = (np.empty(m, dtype=np.uint8), np.empty(m, dtype=np.float32), np.empty(m, dtype=np.float32))
    rfit_out = (pinned.ones((n, K_COMPONENTS), np.float64), np.ones(K_COMPONENTS, dtype=np.float64), None)

    def ingest():
        return ctx.ingest_bed(host.ptr, n, m, qc=None, vcf_maf=0.01, out=stats_out)

    keep, mean, sd, _, d_kept = ingest()
    r = timed_rfit(torch, None, ctx, dev, 1, max(args.steps, 10), max(args.warmup, 3), rfit_out)
    passes = 2 * POWER_ITERS + 3
    bytes_per_pass = d_kept * bps
    t_kern = r["sk_kernel_ms"] / max(r["sk_n"], 1) * 1e-3
    rec = {"workload": f"1000G-shape rfit k={K_COMPONENTS} oversample={OVERSAMPLE} q={POWER_ITERS}: {n} samples x {m} SNPs "
                       f"({d_kept} after MAF 0.01) on one GPU",
           "ms_per_step": r["t_step"] * 1e3, "value": passes * bytes_per_pass / r["t_step"] / 1e9, "unit": "GB/s",
           "gpu_launches": int(r["launches"]),
           "roofline": {"bound": "hbm", "kernel": "sketch_i8_kernel", "achieved": bytes_per_pass / t_kern / 1e9 if t_kern else 0.0,
                        "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": bytes_per_pass / t_kern / 1e9 / pk["hbm_gbs"] if t_kern else 0.0,
                        "ms_per_launch": t_kern * 1e3, "peak_source": pk_src,
                        "kernel_share_of_step": (r["sk_kernel_ms"] * 1e-3 / max(args.steps, 10)) / r["t_step"]},
           "eigenvalues_head": [float(x) for x in r["eigenvalues"][:3]]}
    if not args.no_e2e:
        def e2e_step():
            ingest()
            ctx.rfit(K_COMPONENTS, OVERSAMPLE, power_iters=POWER_ITERS, seed=RFIT_SEED, want_loadings=False, out=rfit_out)
        e2e_step()
        te = e2e_loop(torch, None, dev, 1, e2e_step, 3)
        rec["e2e"] = {"value": passes * bytes_per_pass / te / 1e9, "unit": "GB/s", "ms_per_step": te * 1e3, "steps": 3,
                      "h2d_bytes_per_step": int(m * bps), "d2h_bytes_per_step": int(n * K_COMPONENTS * 8 + m * 16)}
    if not args.no_parity:
        sc, ev, _ = ctx.rfit(K_COMPONENTS, OVERSAMPLE, power_iters=POWER_ITERS, seed=RFIT_SEED, want_loadings=False)
        sc5, ev5, _ = ctx.rfit(K_COMPONENTS, OVERSAMPLE, power_iters=5, seed=RFIT_SEED, want_loadings=False)
        ctx.close()
        ctx = None
        t0 = time.perf_counter()
        payload = torch.empty((m, bps), dtype=torch.uint8, device=dev)
        gen = gp.Context(dev.index or 0)
        gen.synth_bed_device(payload.data_ptr(), n, m, 0, DATA_SEED, N_POPS, FST, 0.0, FST_GRADE)
        # K-a alone (the decode / allele-count kernel of the ingest) on the resident payload: HBM-bound
        ka_ms = gen.count_kernel_ms(payload.data_ptr(), n, m - 1, reps=10)      # (m - 1 rows: 16 readable bytes past the end)
        rec["ka_counts_kernel"] = {"ms_per_launch": ka_ms, "bytes": int((m - 1) * bps), "achieved_GBps": (m - 1) * bps / ka_ms / 1e6,
                                   "frac_of_hbm_peak": (m - 1) * bps / ka_ms / 1e6 / pk["hbm_gbs"],
                                   "note": "chunk_counts_kernel on 626-byte rows at the file pitch (unaligned rows)"}
        gen.close()
        ev_x, v_x = exact_pca_f64(torch, payload, n, np.asarray(keep[:m], dtype=bool), mean[:m], sd[:m], K_COMPONENTS + 1)
        del payload
        torch.cuda.empty_cache()
        k = K_COMPONENTS
        rec["parity_at_scale"] = {
            "against": "exact f64 eigen-decomposition of the N x N Gram matrix of the standardized matrix (torch f64, "
                       "chunked over the decoded .bed payload; no library kernel on that side)",
            "shape": f"{n} x {d_kept}", "tolerance": "1e-4 relative / 1e-3 rad",
            "converged_q5": {"eigenvalue_max_rel_err": float(np.max(np.abs(ev5 / ev_x[:k] - 1.0))),
                             "score_subspace_angle_rad": subspace_angle(sc5, v_x[:, :k])},
            "benchmark_q%d" % POWER_ITERS: {"eigenvalue_max_rel_err": float(np.max(np.abs(ev / ev_x[:k] - 1.0))),
                                            "score_subspace_angle_rad": subspace_angle(sc, v_x[:, :k]),
                                            "leading_8_eigenvalue_max_rel_err": float(np.max(np.abs(ev[:8] / ev_x[:8] - 1.0))),
                                            "note": "q = 2 is the reference's default: the randomized method itself has not "
                                                    "converged on the trailing components (Ritz values approach from below, "
                                                    "geometrically in q); the converged run is the arithmetic check"},
            "eigenvalue_gap_k_to_k1": float(ev_x[k - 1] / ev_x[k]), "seconds": time.perf_counter() - t0}
    if ctx is None:
        ctx = gp.Context(dev.index or 0)      # (only to give the pinned buffers back)
        pinned.ctx = host.ctx = ctx
    pinned.free()
    host.free()
    ctx.close()
    return rec


def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist
    import genomic_pca_b200 as gp

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk, pk_src = peaks()
    n, m_total, nb_total = args.samples, args.snps, args.blocks
    bps = (n + 3) // 4
    b0, b1, s0, s1 = shard_of(rank, world, nb_total, m_total)
    m = s1 - s0
    nb = b1 - b0
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    host_threads = max(1, cores // max(local_world, 1))

    ctx = gp.Context(local_rank)
    numa_cpus = ctx.bind_host_to_device() if local_world > 1 else 0     # payload pages NUMA-local to the rank's GPU
    if args.engine is not None:
        ctx.set_sketch_engine(args.engine)
    ctx.set_host_threads(host_threads)
    ctx.set_sketch_timing(True)
    if world > 1:
        # the library's own NCCL communicator: rank 0 creates the id, torch.distributed only carries the bytes
        box = [gp.binding.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0], rank, world)
        ctx.set_shard(s0, m_total)
    host = HostPayload(ctx, n, m, s0)
    pinned = PinnedArrays(ctx)

    # ---- rfit: resident step time (value) and e2e ------------------------------------------------
    stats_out = (np.empty(m, dtype=np.uint8), np.empty(m, dtype=np.float32), np.empty(m, dtype=np.float32),
                 np.empty(m, dtype=np.uint8))
    rfit_out = (pinned.ones((n, K_COMPONENTS), np.float64), np.ones(K_COMPONENTS, dtype=np.float64), None)
    t0 = time.perf_counter()
    keep, mean, sd, _, d_kept = ctx.ingest_bed(host.ptr, n, m, qc=None, vcf_maf=0.01, out=stats_out[:3])
    t_first_ingest = time.perf_counter() - t0
    resident_rfit = ctx.resident_snp_rows
    passes = 2 * POWER_ITERS + 3
    bytes_per_pass = d_kept * bps                          # this rank's algorithmic bytes: M_loc * ceil(N/4)
    job_bytes_per_pass, = sum_over_ranks(torch, dist, dev, world, [float(bytes_per_pass)])
    d_total, = sum_over_ranks(torch, dist, dev, world, [float(d_kept)])
    # the scores are the same on every shard: rank 0's copy goes to the host (what the CLI writes), the others pass NULL
    want_sc = rank == 0
    r = timed_rfit(torch, dist, ctx, dev, world, args.steps, args.warmup, rfit_out, ClockSampler(local_rank), want_sc)
    t_step = r["t_step"]
    value = passes * job_bytes_per_pass / t_step / 1e9
    e2e = None
    if not args.no_e2e:
        def e2e_rfit():
            ctx.ingest_bed(host.ptr, n, m, qc=None, vcf_maf=0.01, out=stats_out[:3])
            ctx.rfit(K_COMPONENTS, OVERSAMPLE, power_iters=POWER_ITERS, seed=RFIT_SEED, want_loadings=False, out=rfit_out,
                     want_scores=want_sc)
        te = e2e_loop(torch, dist, dev, world, e2e_rfit, args.e2e_steps)
        e2e = {"value": passes * job_bytes_per_pass / te / 1e9, "unit": "GB/s", "ms_per_step": te * 1e3,
               "steps": args.e2e_steps, "h2d_bytes_per_step": int(m_total * bps),
               "d2h_bytes_per_step": int(n * K_COMPONENTS * 4 + world * K_COMPONENTS * 8 + m_total * 16),
               "pca_wall_s": te, "includes": "gpca_ingest_bed of every rank's shard from pinned host memory + gpca_rfit + "
                                             "scores / eigenvalues on the host"}

    # ---- the kernel's arithmetic at this size, on sampled output rows --------------------------------------------
    # One-GPU runs only: on a sharded context gpca_sketch_sample_side is a COLLECTIVE (it ends with the exchange of the
    # N x l sketch), so a check issued by one rank alone would leave that rank waiting for the others for ever.
    parity_full = None
    if world == 1 and not args.no_parity:
        try:
            view = np.ctypeslib.as_array(ctypes.cast(host.ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(m * bps,)).reshape(m, bps)
            parity_full = sampled_parity(torch, ctx, view, n, keep, mean, sd, dev)
            parity_full["shape"] = f"{n} samples x {int(d_kept)} SNPs"
            del view
        except Exception as e:      # supplementary record
            parity_full = {"error": repr(e)}
        torch.cuda.empty_cache()

    # ---- EigenSNP on the same configuration --------------------------------------------------------
    es = None
    if not args.no_eigensnp:
        try:
            cfg = gp.EigenSnpConfig(target_num_global_pcs=K_COMPONENTS)
            ctx.set_memory_reserve(gp.binding.eigensnp_workspace_bytes(n, m, nb, cfg))
            qc = gp.QcConfig(0.98, 0.01, 1.0)                 # HWE off: pooled structured populations fail it

            def es_ingest():
                return ctx.ingest_bed(host.ptr, n, m, qc=qc, out=stats_out)

            _, _, _, _, d_es = es_ingest()
            edges = np.linspace(0, d_es, nb + 1).astype(np.int64)
            blocks = [np.arange(edges[i], edges[i + 1], dtype=np.uint64) for i in range(nb)]
            es_out = (pinned.ones((n, K_COMPONENTS), np.float32), np.ones(K_COMPONENTS, dtype=np.float64),
                      pinned.ones((d_es, K_COMPONENTS), np.float32))
            ctx.eigensnp(blocks, cfg, out=es_out, want_scores=want_sc)
            reps = 3
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ctx.reset_launch_count()
            t0 = time.perf_counter()
            for _ in range(reps):
                sc_es, ev_es, _ = ctx.eigensnp(blocks, cfg, out=es_out, want_scores=want_sc)
            t_es = (time.perf_counter() - t0) / reps
            es_launches = ctx.launch_count // reps
            t_es, = max_over_ranks(torch, dist, dev, world, [t_es])
            es = {"workload": f"EigenSNP k={K_COMPONENTS}, {nb_total} LD blocks, effective defaults of src/main.rs:545-588",
                  "resident_wall_s": t_es, "gpu_launches": int(es_launches), "resident_snp_rows_frac": ctx.resident_snp_rows / max(d_es, 1),
                  "eigenvalues_head": [float(x) for x in ev_es[:3]]}
            if not args.no_e2e:
                def e2e_es():
                    es_ingest()
                    ctx.eigensnp(blocks, cfg, out=es_out, want_scores=want_sc)
                te_es = e2e_loop(torch, dist, dev, world, e2e_es, args.e2e_steps)
                es["e2e"] = {"pca_wall_s": te_es, "ms_per_step": te_es * 1e3, "steps": args.e2e_steps,
                             "h2d_bytes_per_step": int(m_total * bps),
                             "d2h_bytes_per_step": int(n * K_COMPONENTS * 4 + m_total * (16 + K_COMPONENTS * 4))}
        except Exception as e:       # the headline record must not be lost to the second mode
            es = {"error": repr(e)}
            if world > 1:
                raise
    pinned.free()
    host.free()
    ctx.close()
    torch.cuda.empty_cache()

    c3 = None
    if not args.no_c3 and rank == 0:
        try:
            c3 = run_c3(args, torch, gp, dev, pk, pk_src)
        except Exception as e:  # the headline record must not be lost to a failure of the supplementary one
            c3 = {"error": repr(e)}
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # dominant kernel = the sketch kernel; its launches are bracketed by CUDA events on the library's stream
    # (one launch per pass and resident segment); the pass-level figure also contains operand prep and the split-K reduce.
    t_kern = (r["sk_kernel_ms"] / max(r["sk_n"], 1)) * 1e-3
    t_pass = (r["sk_ms"] / max(r["sk_n"], 1)) * 1e-3
    bpl = bytes_per_pass * passes * args.steps / max(r["sk_n"], 1)      # bytes per launch (== bytes_per_pass when unsegmented)
    ach_gbs = bpl / t_kern / 1e9 if t_kern > 0 else 0.0
    flops_per_launch = 2.0 * 4.0 * bpl * (K_COMPONENTS + OVERSAMPLE)
    ach_tf = flops_per_launch / t_kern / 1e12 if t_kern > 0 else 0.0
    l = K_COMPONENTS + OVERSAMPLE
    eng = r["engine"]          # the engine that actually ran (the library falls back by shape / l)
    kernel_name = {0: "sketch_simt_kernel", 1: "sketch_tc_kernel", 2: "sketch_i8_kernel" if l <= 32 else "sketch_tc_kernel<64>"}[eng]
    roofline = {"bound": "hbm", "achieved": ach_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach_gbs / pk["hbm_gbs"],
                "traffic": 1.007 * bpl,
                "traffic_note": "algorithmic bytes x 1.007 = the DRAM-to-algorithmic ratio of the committed ncu --set full capture of "
                                "this kernel (profiles/r2_ncu_full_sketch_i8_kernel_shard_summary.csv: 10.950 GB read + 0.064 GB "
                                "written for 10.9375 GB of packed rows); not re-measured in this run",
                "peak_source": pk_src, "kernel": kernel_name, "ms_per_launch": t_kern * 1e3,
                "launches_per_step": r["sk_n"] / args.steps, "bytes_per_launch": bpl,
                "kernel_share_of_step": (r["sk_kernel_ms"] * 1e-3 / args.steps) / t_step,
                "pass_level": {"ms_per_pass": t_pass * 1e3, "achieved": bpl / t_pass / 1e9 if t_pass > 0 else 0.0,
                               "frac": (bpl / t_pass / 1e9) / pk["hbm_gbs"] if t_pass > 0 else 0.0,
                               "note": "operand prep + sketch kernel + split-K reduce"},
                "tensor_achieved_tflops": ach_tf, "tensor_frac_of_sustained_bf16": ach_tf / pk["bf16_tflops_sustained"]}
    line = {
        "metric": "genotype_GBps_per_sketch_pass", "value": value, "unit": "GB/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": {0: "f32", 1: "f16 x f16 -> f32 (tcgen05)", 2: "u8 x s8 -> s32 (tcgen05 kind::i8, exact)"}[eng],
        "data": "synthetic",
        "config": {"workload": f"{'BASELINE config 4: ' if (n, m_total) == (C4_SAMPLES, C4_SNPS) else ''}rfit k={K_COMPONENTS} "
                               f"oversample={OVERSAMPLE} q={POWER_ITERS} on {n} samples x {m_total} SNPs ({int(d_total)} after MAF 0.01), "
                               f"2-bit packed, SNPs / {nb_total} LD blocks split over {world} GPU(s)",
                   "passes_per_step": passes, "bytes_per_pass": job_bytes_per_pass,
                   "l2": "inputs larger than L2 (packed shard >> 126 MB)" if bytes_per_pass > 2e8 else "input fits L2",
                   "pca_wall_s": t_step, "host_wall_s_per_step": r["wall_per_step"], "host_threads_per_rank": host_threads,
                   "numa_local_cpus_rank0": numa_cpus,
                   "result_buffers": "caller-owned pinned host memory (gpca_host_alloc); scores requested on rank 0 only",
                   "exchange": "ncclAllReduce issued by the library (gpca_comm_init)" if world > 1 else "none (one shard)",
                   "collectives_per_step": r["collectives_per_step"],
                   "resident_snp_rows_frac_rank0": resident_rfit / max(d_kept, 1),
                   "payload_setup_s": {"pinned_alloc": host.alloc_s, "synthesise": host.fill_s, "first_ingest": t_first_ingest}},
        "clocks": r["clocks"], "e2e": e2e, "gpu_launches": int(r["launches"]), "roofline": roofline,
        "eigenvalues_head": [float(x) for x in r["eigenvalues"][:3]],
    }
    if parity_full is not None:
        line["parity_full_size"] = parity_full
    if es is not None:
        line["eigensnp"] = es
    if c3 is not None:
        line["c3"] = c3
        if isinstance(c3, dict) and "parity_at_scale" in c3:
            line["parity_at_scale"] = c3["parity_at_scale"]
    if not args.no_cpu:
        lim = blas_threads(cores)
        dt, cp, cb = cpu_rfit_sample(n, args.cpu_snps)
        line["cpu_baseline"] = {"value": cp * cb / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
                                "blas_threads": cores if lim is not None else os.environ.get("OMP_NUM_THREADS", "default"),
                                "sample": f"{n} samples x {args.cpu_snps} SNPs, one rfit "
                                f"(numpy/OpenBLAS f64 restatement of the reference's CPU path), {dt:.1f} s"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def start_watchdog(limit_s=None):
    """A run of this script takes one to three minutes.  If it is still going after `limit_s` seconds (default 900,
    BENCH_WATCHDOG_S; 0 = off) something is waiting that will never come -- e.g. a collective that one rank did not
    issue -- and the process ends itself with a one-line explanation instead of holding the GPUs until an outer limit."""
    if limit_s is None:
        limit_s = float(os.environ.get("BENCH_WATCHDOG_S", "900"))
    if limit_s <= 0:
        return None

    def fire():
        sys.stderr.write(f"bench.py: no result after {limit_s:.0f} s (rank {os.environ.get('RANK', '0')}): giving up\n")
        sys.stderr.flush()
        os._exit(3)

    t = threading.Timer(limit_s, fire)
    t.daemon = True
    t.start()
    return t


def main():
    args = parse_args()
    start_watchdog()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        print(json.dumps({"error": "launch with torchrun for --gpus > 1"}))
        sys.exit(2)
    run_ours(args, rank, world)


if __name__ == "__main__":
    main()
