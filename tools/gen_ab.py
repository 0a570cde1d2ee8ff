"""Same-box A/B of the generated-operand first pass (development aid)."""
import os, sys, subprocess
for rnd in range(2):
    for env in ({}, {"GPCA_DEBUG_NO_GEN_FUSE": "1"}):
        e = dict(os.environ); e.update(env); e["AB_SHAPES"] = "c3"
        p = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ab_time.py"), "--child"],
                           capture_output=True, text=True, env=e)
        res = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        print(sorted(env) or "fused", res[0][7:] if res else p.stderr[-300:], flush=True)
