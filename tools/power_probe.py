"""Power / clock behaviour of the sketch pass under sustained load (run on the GPU box).

  python tools/power_probe.py [seconds]

Runs rfit back to back for a few seconds at the 500,000 x 87,500 shard shape and at 2,504 x 10M while NVML is sampled
every 10 ms: SM clock, power draw, throttle reasons.  The kernel's own cycle counter (GPCA_I8_PROF builds) showed the SMs
running at ~1.3 GHz inside the sketch kernel; this probe shows whether that is the power cap.
"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
    import pynvml
    import torch
    import bench
    import genomic_pca_b200 as gp
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    print("power limit W:", pynvml.nvmlDeviceGetPowerManagementLimit(h) / 1000.0,
          "max sm MHz:", pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM), flush=True)
    for (n, m) in [(500_000, 87_500), (2504, 10_000_000)]:
        payload = bench.synth_bed_device(torch, n, m, 0, dev)
        ctx = gp.Context(0)
        ctx.load_bed_device(payload.data_ptr(), n, m)
        keep, mean, sd = ctx.vcf_maf_filter(0.01)
        d = ctx.set_pca_snps_mask(keep, mean, sd)
        del payload
        torch.cuda.empty_cache()
        for _ in range(2):
            ctx.rfit(20, 10, power_iters=2, seed=42, want_loadings=False)
        torch.cuda.synchronize()
        rows = []
        stop = [False]

        def sample():
            while not stop[0]:
                try:
                    rows.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                                 pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM),
                                 pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                 pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h),
                                 pynvml.nvmlDeviceGetTemperature(h, pynvml.NVML_TEMPERATURE_GPU)))
                except Exception as e:      # noqa
                    rows.append((time.perf_counter(), -1, -1, -1.0, -1, -1))
                time.sleep(0.01)

        th = threading.Thread(target=sample, daemon=True)
        th.start()
        time.sleep(0.3)
        t_idle_end = time.perf_counter()
        ctx.sketch_stats(reset=True)
        t0 = time.perf_counter()
        steps = 0
        while time.perf_counter() - t0 < secs:
            ctx.rfit(20, 10, power_iters=2, seed=42, want_loadings=False)
            steps += 1
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        sk_ms, _, sk_n = ctx.sketch_stats(reset=True)
        kern = ctx.last_kernel_ms / max(sk_n, 1)
        time.sleep(0.3)
        stop[0] = True
        th.join()
        load = [r for r in rows if t0 + 0.5 < r[0] < t1]
        idle = [r for r in rows if r[0] < t_idle_end]
        med = lambda xs: sorted(xs)[len(xs) // 2] if xs else None      # noqa
        reasons = set()
        for r in load:
            reasons.add(r[4])
        print(json.dumps({"shape": f"{n}x{m}", "steps": steps, "step_ms": (t1 - t0) / steps * 1e3, "kernel_ms": kern,
                          "sm_mhz_load_median": med([r[1] for r in load]), "sm_mhz_load_min": min(r[1] for r in load),
                          "sm_mhz_load_max": max(r[1] for r in load), "mem_mhz": med([r[2] for r in load]),
                          "power_w_load_median": med([r[3] for r in load]), "power_w_load_max": max(r[3] for r in load),
                          "power_w_idle": med([r[3] for r in idle]), "temp_c_max": max(r[5] for r in load),
                          "throttle_reason_bitmasks": sorted(hex(x) for x in reasons), "samples": len(load)}), flush=True)
        ctx.close()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
