#!/usr/bin/env python
"""bench.py -- the hot path (randomized-SVD sketch passes of the rfit PCA) on synthetic data.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--snps M] [--samples N]

One "step" = one complete rfit PCA (1 + 2q + 2 sketch passes over the resident 2-bit matrix, the
re-orthonormalisations and the small eigensolves) on the workload named in config.workload.
metric = genotype GB/s per sketch pass = packed genotype bytes streamed by the sketch passes / time.
  value : inputs resident in HBM when the timed region starts (device-timed, CUDA events, max over ranks)
  e2e   : the same PCA through the C ABI from HOST buffers (pinned .bed payload -> H2D -> counts ->
          QC -> resident build -> rfit -> scores back on the host), H2D/D2H inside the timed region.
N > 1 (torchrun): SNP-sharded, one rank per GPU, weak scaling (per-GPU shard fixed); the N x l sketch
is summed over ranks with NCCL after every sample-side pass (the path's one exchange step).
"""
import argparse
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


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=None, help="timed steps (default 20; 3 for --impl reference)")
    p.add_argument("--warmup", type=int, default=None, help="untimed steps (default 5; 1 for --impl reference)")
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--samples", type=int, default=2504)
    p.add_argument("--snps", type=int, default=10_000_000, help="SNPs per GPU (weak scaling)")
    p.add_argument("--engine", type=int, default=None)
    p.add_argument("--components", type=int, default=None, help="k (default 20 = BASELINE config 3; config 5 uses 40)")
    p.add_argument("--power-iters", type=int, default=None, help="q (default 2; config 5 uses 4)")
    p.add_argument("--cpu-snps", type=int, default=150_000, help="SNP rows of the CPU baseline sample")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-ukb", action="store_true", help="skip the supplementary 500k x 87.5k shard measurement")
    a = p.parse_args()
    # defaults: a timed region long enough (~0.3 s) for the clock sampler to see the run under load; the CPU arm's steps
    # are seconds each
    if a.steps is None:
        a.steps = 3 if a.impl == "reference" else 20
    if a.warmup is None:
        a.warmup = 1 if a.impl == "reference" else 5
    global K_COMPONENTS, POWER_ITERS
    if a.components is not None:
        K_COMPONENTS = a.components
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
def synth_bed_device(torch, n_samples, n_snps, snp_offset, device):
    """Balding-Nichols genotypes generated on the device straight into PLINK .bed layout
    (P = k+2 populations, ancestral AF ~ U(0.05,0.5), F_ST = 0.1, no missing calls) by the library's
    counter-based generator (Philox keyed by (DATA_SEED, global SNP index, sample)): rows [a, b) of any shard are the
    rows [a, b) of the whole matrix, whatever the number of GPUs."""
    import genomic_pca_b200 as gp
    bps = (n_samples + 3) // 4
    out = torch.empty((n_snps, bps), dtype=torch.uint8, device=device)
    gen = gp.Context(device.index or 0)
    gen.synth_bed_device(out.data_ptr(), n_samples, n_snps, snp_offset, DATA_SEED, N_POPS, 0.1, 0.0)
    gen.close()
    return out


def cpu_rfit_sample(n_samples, n_snps, steps=1):
    """The reference's CPU algorithm (oracle restatement: f64 matrix, OpenBLAS GEMMs, numpy QR/eigh)
    timed on a bounded sample of the same workload.  Returns (seconds per step, passes, bytes per pass)."""
    from oracle import pca, synth
    g = np.concatenate([synth.balding_nichols(n_samples, min(20000, n_snps - c0), n_pops=N_POPS, seed=7 + c0)[0]
                        for c0 in range(0, n_snps, 20000)])
    mean = g.mean(1)
    sd = g.std(1, ddof=1)
    ok = sd > 1e-9
    g = g[ok]
    t0 = time.perf_counter()
    for _ in range(steps):
        # build_matrix (vcf.rs:317-345: u8 -> f64) + standardise + rfit + transform, as the reference does per run
        S = pca.standardize_dense(g, mean[ok], sd[ok])
        pca.rfit(S, K_COMPONENTS, OVERSAMPLE, seed=RFIT_SEED, power_iters=POWER_ITERS)
    dt = (time.perf_counter() - t0) / steps
    passes = 2 * POWER_ITERS + 3
    return dt, passes, g.shape[0] * ((n_samples + 3) // 4)


# ------------------------------------------------------------------------------------------
def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = []
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_rfit_sample(args.samples, args.cpu_snps)
    for _ in range(args.steps):
        dt, passes, bpp = cpu_rfit_sample(args.samples, args.cpu_snps)
        per_step.append(dt)
    dt = float(np.mean(per_step))
    val = passes * bpp / dt / 1e9
    line = {
        "impl": "reference", "metric": "genotype_GBps_per_sketch_pass", "value": val, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"rfit k={K_COMPONENTS} oversample={OVERSAMPLE} q={POWER_ITERS}, "
                               f"{args.samples} samples x {args.cpu_snps} SNPs (bounded CPU sample of the "
                               f"1000G-shape config: {args.samples} x {args.snps})",
                   "passes_per_step": passes},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port",
                         "sample": f"{args.samples} samples x {args.cpu_snps} SNPs, numpy/OpenBLAS f64 restatement "
                                   "(reference binary cannot be built here: no cargo/rustc)"},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def make_allreduce_hook(torch, dist, dev):
    """The library's one exchange step (sum of a device buffer over the shards) bound to NCCL through torch.distributed,
    on the library's own stream."""
    def hook(ptr, count, dtype, stream):
        es = torch.cuda.ExternalStream(stream, device=dev)
        td = torch.float32 if dtype == 0 else torch.float64
        iface = {"shape": (count,), "typestr": "<f4" if dtype == 0 else "<f8", "data": (ptr, False), "version": 2}
        holder = type("P", (), {"__cuda_array_interface__": iface})()
        t = torch.as_tensor(holder, device=dev)
        assert t.dtype == td
        with torch.cuda.stream(es):
            dist.all_reduce(t)
    return hook


def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist
    import genomic_pca_b200 as gp

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n, m = args.samples, args.snps
    bps = (n + 3) // 4
    payload = synth_bed_device(torch, n, m, rank * m, dev)
    torch.cuda.synchronize()

    ctx = gp.Context(local_rank)
    if args.engine is not None:
        ctx.set_sketch_engine(args.engine)
    if world > 1:
        ctx.set_allreduce(make_allreduce_hook(torch, dist, dev))
        ctx.set_shard(rank * m, world * m)

    def prepare_from_device():
        ctx.load_bed_device(payload.data_ptr(), n, m)
        keep, mean, sd = ctx.vcf_maf_filter(0.01)
        return ctx.set_pca_snps_mask(keep, mean, sd)

    d_kept = prepare_from_device()
    passes = 2 * POWER_ITERS + 3
    bytes_per_pass = d_kept * bps                          # algorithmic: M_loc * ceil(N/4)
    flops_per_pass = 2.0 * n * d_kept * (K_COMPONENTS + OVERSAMPLE)

    # caller-owned result buffers, allocated (and touched) once, as a host application would
    rfit_out = (np.ones((n, K_COMPONENTS), dtype=np.float64), np.ones(K_COMPONENTS, dtype=np.float64), None)

    def step():
        return ctx.rfit(K_COMPONENTS, OVERSAMPLE, power_iters=POWER_ITERS, seed=RFIT_SEED, want_loadings=False,
                        out=rfit_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    ctx.sketch_stats(reset=True)
    ctx.reset_launch_count()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    lib_stream = torch.cuda.ExternalStream(ctx.stream, device=dev)      # the stream the kernels are launched on
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(lib_stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sc, ev, _ = step()
    ev1.record(lib_stream)
    barrier()
    wall = time.perf_counter() - t0
    dev_s = ev0.elapsed_time(ev1) * 1e-3
    clocks = sampler.stop()
    launches = ctx.launch_count
    sk_ms, sk_bytes, sk_n = ctx.sketch_stats(reset=True)
    sk_kernel_ms = ctx.last_kernel_ms
    # device time between CUDA events on the library's stream (the wall clock is kept as a cross-check)
    t_step = torch.tensor([dev_s / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_step, op=dist.ReduceOp.MAX)
    t_step = float(t_step.item())
    value = world * passes * bytes_per_pass / t_step / 1e9

    # ---- e2e through the C ABI from pinned host memory --------------------------------------
    e2e = None
    if not args.no_e2e:
        host = torch.empty((m, bps), dtype=torch.uint8, pin_memory=True)
        host.copy_(payload)
        torch.cuda.synchronize()
        del payload
        e2e_steps = max(1, min(args.steps, 3))

        stats_out = (np.empty(m, dtype=np.uint8), np.empty(m, dtype=np.float32), np.empty(m, dtype=np.float32))

        def e2e_step():
            # one streaming ingest call (H2D copy, counts, MAF filter on host threads, resident matrices), then rfit;
            # the keep mask and the per-SNP mean / sd come back to the host as in the reference's VCF flow
            keep, mean, sd, _, d_pca = ctx.ingest_bed(host.data_ptr(), n, m, qc=None, vcf_maf=0.01, out=stats_out)
            return ctx.rfit(K_COMPONENTS, OVERSAMPLE, power_iters=POWER_ITERS, seed=RFIT_SEED, want_loadings=False,
                            out=rfit_out)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            sc2, ev2, _ = e2e_step()
        barrier()
        te = (time.perf_counter() - t0) / e2e_steps
        te_t = torch.tensor([te], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te_t, op=dist.ReduceOp.MAX)
        te = float(te_t.item())
        e2e = {"value": world * passes * bytes_per_pass / te / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": int(m * bps), "d2h_bytes_per_step": int(n * K_COMPONENTS * 8 + K_COMPONENTS * 8 + m * 16),
               "ms_per_step": te * 1e3, "steps": e2e_steps}

    # supplementary measurement at the shape the headline metric is quoted on (all ranks take part: on N GPUs it is
    # the 500k-sample array with N x 87.5k SNPs, EigenSNP with the N x l exchange over NCCL)
    ukb = None
    if not args.no_ukb and (n, m) == (2504, 10_000_000):
        ctx.close()
        if not args.no_e2e:
            del host
        else:
            del payload
        torch.cuda.empty_cache()
        ukb = ukb_shard_supplement(torch, gp, dev, peaks()[0], rank, world, dist if world > 1 else None)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    # dominant kernel = the sketch kernel; its launches are bracketed by CUDA events on the library's stream
    # (one launch per pass); the pass-level figure also contains operand prep and the split-K reduce.
    t_kern = (sk_kernel_ms / max(sk_n, 1)) * 1e-3
    t_pass = (sk_ms / max(sk_n, 1)) * 1e-3
    ach_gbs = bytes_per_pass / t_kern / 1e9 if t_kern > 0 else 0.0
    ach_tf = flops_per_pass / t_kern / 1e12 if t_kern > 0 else 0.0
    eng = ctx_engine(ctx, args)
    roofline = {"bound": "hbm", "achieved": ach_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach_gbs / pk["hbm_gbs"], "traffic": TRAFFIC_NCU.get(eng) if (n, m) == (2504, 10_000_000) else None, "peak_source": pk_src,
                "kernel": {0: "sketch_simt_kernel", 1: "sketch_tc_kernel",
                           2: "sketch_i8_kernel" if K_COMPONENTS + OVERSAMPLE <= 32 else "sketch_tc_kernel<64>"}[eng],
                "ms_per_launch": t_kern * 1e3, "launches_per_step": passes,
                "kernel_share_of_step": (sk_kernel_ms * 1e-3 / args.steps) / t_step,
                "pass_level": {"ms_per_pass": t_pass * 1e3, "achieved": bytes_per_pass / t_pass / 1e9 if t_pass > 0 else 0.0,
                               "frac": (bytes_per_pass / t_pass / 1e9) / pk["hbm_gbs"] if t_pass > 0 else 0.0,
                               "note": "operand prep + sketch kernel + split-K reduce"},
                "tensor_achieved_tflops": ach_tf, "tensor_frac_of_sustained_bf16": ach_tf / pk["bf16_tflops_sustained"]}
    line = {
        "metric": "genotype_GBps_per_sketch_pass", "value": value, "unit": "GB/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": {0: "f32", 1: "f16 x f16 -> f32 (tcgen05)", 2: "u8 x s8 -> s32 (tcgen05 kind::i8, exact)"}[ctx_engine(ctx, args)],
        "data": "synthetic",
        "config": {"workload": f"{'1000G-shape ' if (n, m) == (2504, 10_000_000) else ''}rfit k={K_COMPONENTS} oversample={OVERSAMPLE} q={POWER_ITERS}: "
                               f"{n} samples x {m} SNPs per GPU ({d_kept} after MAF 0.01), 2-bit packed, SNP-sharded",
                   "passes_per_step": passes, "bytes_per_pass": bytes_per_pass,
                   "l2": "inputs larger than L2 (packed shard >> 126 MB)" if bytes_per_pass > 2e8 else "input fits L2",
                   "pca_wall_s": t_step, "host_wall_s_per_step": wall / args.steps},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "eigenvalues_head": [float(x) for x in ev[:3]],
    }
    if ukb is not None:
        line["ukb_shard"] = ukb
    if not args.no_cpu:
        dt, cp, cb = cpu_rfit_sample(n, args.cpu_snps)
        line["cpu_baseline"] = {"value": cp * cb / dt / 1e9, "unit": "GB/s", "cores": os.cpu_count() or 1,
                                "kind": "port", "sample": f"{n} samples x {args.cpu_snps} SNPs, one rfit "
                                f"(numpy/OpenBLAS f64 restatement of the reference's CPU path), {dt:.1f} s"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, bytes per launch) of sketch_i8_kernel at config 3 from the
# committed ncu capture (profiles/r1_ncu_full_sketch_i8_kernel_c3_final2.csv): sample side 7.378 + 0.040 GB, snp side
# 6.480 + 1.268 GB; one rfit runs 4 sample-side and 3 snp-side launches.  Only valid for the default workload.
TRAFFIC_NCU = {2: (4 * 7.418e9 + 3 * 7.748e9) / 7}


def ukb_shard_supplement(torch, gp, dev, pk, rank=0, world=1, dist=None):
    """Supplementary measurement at the shape the headline metric is quoted on: the per-GPU shard of BASELINE
    config 4 on 8 GPUs (500,000 samples x 87,500 SNPs).  The full config needs >= 2 GPUs in this round's
    two-orientation layout, so on one GPU its shard is measured: sketch-pass roofline fraction + EigenSNP wall time."""
    n, m, nblocks = 500_000, 87_500, 212
    payload = synth_bed_device(torch, n, m, rank * m, dev)
    torch.cuda.synchronize()
    ctx = gp.Context(dev.index or 0)
    if world > 1:      # every rank holds one shard of SNPs / LD blocks of the same 500k samples
        ctx.set_allreduce(make_allreduce_hook(torch, dist, dev))
        ctx.set_shard(rank * m, world * m)
    ctx.load_bed_device(payload.data_ptr(), n, m)
    keep, mean, sd, _ = ctx.snp_qc(gp.QcConfig(0.98, 0.01, 1.0))     # HWE off: pooled structured populations fail it
    d = ctx.set_pca_snps_mask(keep, mean, sd)
    del payload
    torch.cuda.empty_cache()
    bps = (n + 3) // 4
    # caller-owned result buffers, allocated and touched once (see Context.rfit)
    rfit_out = (np.ones((n, K_COMPONENTS), dtype=np.float64), np.ones(K_COMPONENTS, dtype=np.float64), None)
    for _ in range(2):
        ctx.rfit(K_COMPONENTS, OVERSAMPLE, power_iters=POWER_ITERS, seed=RFIT_SEED, want_loadings=False, out=rfit_out)
    ctx.sketch_stats(reset=True)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        ctx.rfit(K_COMPONENTS, OVERSAMPLE, power_iters=POWER_ITERS, seed=RFIT_SEED, want_loadings=False, out=rfit_out)
    t_rfit = (time.perf_counter() - t0) / reps
    sk_ms, _, sk_n = ctx.sketch_stats(reset=True)
    t_kern = ctx.last_kernel_ms / max(sk_n, 1) * 1e-3
    edges = np.linspace(0, d, nblocks + 1).astype(np.int64)
    blocks = [np.arange(edges[i], edges[i + 1], dtype=np.uint64) for i in range(nblocks)]
    cfg = gp.EigenSnpConfig(target_num_global_pcs=K_COMPONENTS)
    es_out = (np.ones((n, K_COMPONENTS), dtype=np.float32), np.ones(K_COMPONENTS, dtype=np.float64),
              np.ones((d, K_COMPONENTS), dtype=np.float32))
    ctx.eigensnp(blocks, cfg, out=es_out)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    sc, ev, load = ctx.eigensnp(blocks, cfg, out=es_out)
    t_es = time.perf_counter() - t0
    if world > 1:      # max over ranks
        tt = torch.tensor([t_es, t_rfit], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_es, t_rfit = float(tt[0].item()), float(tt[1].item())
    ctx.close()
    gbs = d * bps / t_kern / 1e9 if t_kern > 0 else 0.0
    shape = (f"{n} samples x {m} SNPs on one GPU = the per-GPU shard of BASELINE config 4 (500k x 700k) on 8 GPUs"
             if world == 1 else
             f"{n} samples x {world * m} SNPs ({world * nblocks} LD blocks) sharded over {world} GPUs"
             + (" = BASELINE config 4" if world == 8 else ""))
    return {"shape": shape, "n_gpus": world,
            "sketch_kernel_ms_per_launch": t_kern * 1e3, "sketch_kernel_GBps": gbs,
            "sketch_kernel_frac_of_hbm": gbs / pk["hbm_gbs"], "sketch_pass_ms": sk_ms / max(sk_n, 1),
            "rfit_k20_wall_s": t_rfit, "eigensnp_k20_212_blocks_wall_s": t_es,
            "eigensnp_eigenvalues_head": [float(x) for x in ev[:3]]}


def ctx_engine(ctx, args):
    return args.engine if args.engine is not None else int(os.environ.get("GPCA_SKETCH_ENGINE", "2"))


def main():
    args = parse_args()
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
