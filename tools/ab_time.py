"""Same-box A/B timing of build variants of the sketch engine (run on the GPU box through gpurun).

  python tools/ab_time.py "" "-DGPCA_I8_REGSPLIT=0" ...

For every EXTRA flag set: rebuild libgpca.so (`make EXTRA=...`), then in a fresh process time rfit at BASELINE config 3
(2,504 x 10M) and at the config-4 shard (500,000 x 87,500): ms per rfit step (CUDA events on the library's stream) and
ms per sketch-kernel launch.  Knock-out variants (GPCA_KO_*) give wrong results on purpose; only their timing is read.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "genomic_pca_b200", "csrc")


def child(shapes):
    sys.path.insert(0, ROOT)
    import torch
    import bench
    import genomic_pca_b200 as gp
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    out = {}
    for (n, m) in shapes:
        payload = bench.synth_bed_device(torch, n, m, 0, dev)
        torch.cuda.synchronize()
        ctx = gp.Context(0)
        ctx.set_sketch_timing(True)
        ctx.load_bed_device(payload.data_ptr(), n, m)
        keep, mean, sd = ctx.vcf_maf_filter(0.01)
        d = ctx.set_pca_snps_mask(keep, mean, sd)
        del payload
        torch.cuda.empty_cache()
        for _ in range(3):
            ctx.rfit(20, 10, power_iters=2, seed=42, want_loadings=False)
        ctx.sketch_stats(reset=True)
        st = torch.cuda.ExternalStream(ctx.stream, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        reps = 5
        e0.record(st)
        for _ in range(reps):
            sc, ev, _ = ctx.rfit(20, 10, power_iters=2, seed=42, want_loadings=False)
        e1.record(st)
        torch.cuda.synchronize()
        sk_ms, _, sk_n = ctx.sketch_stats(reset=True)
        kern = ctx.last_kernel_ms / max(sk_n, 1)
        bps = (n + 3) // 4
        out[f"{n}x{m}"] = {"step_ms": round(e0.elapsed_time(e1) / reps, 3), "kernel_ms": round(kern, 4),
                           "pass_ms": round(sk_ms / max(sk_n, 1), 4), "kernel_GBps": round(d * bps / kern / 1e6, 1),
                           "ev0": float(ev[0])}
        if os.environ.get("AB_PROF"):
            # one more step between two markers: the per-warp cycle accounting that a GPCA_I8_PROF build prints
            print(f"MARK begin {n}x{m}", flush=True)
            ctx.rfit(20, 10, power_iters=2, seed=42, want_loadings=False)
            torch.cuda.synchronize()
            print(f"MARK end {n}x{m}", flush=True)
        ctx.close()
        torch.cuda.empty_cache()
    print("RESULT " + json.dumps(out), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        shapes = [(2504, 10_000_000), (500_000, 87_500)]
        if os.environ.get("AB_SHAPES") == "c3":
            shapes = shapes[:1]
        if os.environ.get("AB_SHAPES") == "shard":
            shapes = shapes[1:]
        child(shapes)
        return
    variants = sys.argv[1:] or [""]
    rounds = int(os.environ.get("AB_ROUNDS", "1"))
    for r in range(rounds):
        for v in variants:
            subprocess.check_call(["touch", os.path.join(CSRC, "sketch_i8.cu"), os.path.join(CSRC, "sketch_tc.cu")])
            subprocess.check_call(["make", "-C", CSRC, "-j8", f"EXTRA={v}", "../libgpca.so"], stdout=subprocess.DEVNULL)
            t0 = time.time()
            env = dict(os.environ)
            if "PROF" in v:
                env["AB_PROF"] = "1"
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], capture_output=True, text=True,
                               env=env)
            if "PROF" in v:
                on = False
                for l in p.stdout.splitlines():
                    if l.startswith("MARK begin"):
                        on = True
                    if on and (l.startswith("MARK") or l.startswith("PROF")):
                        print("    " + l)
                    if l.startswith("MARK end"):
                        on = False
            res = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            print(f"[{v or 'default'}] ({time.time() - t0:.0f}s) " + (res[0][7:] if res else "FAILED " + p.stderr[-400:]),
                  flush=True)
    # leave the default build behind
    subprocess.check_call(["touch", os.path.join(CSRC, "sketch_i8.cu"), os.path.join(CSRC, "sketch_tc.cu")])
    subprocess.check_call(["make", "-C", CSRC, "-j8", "../libgpca.so"], stdout=subprocess.DEVNULL)


if __name__ == "__main__":
    main()
