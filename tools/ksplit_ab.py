"""Same-box A/B of the K-split rule (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import subprocess
for rnd in range(2):
    for env in ({}, {"GPCA_DEBUG_OLD_KSPLIT": "1"}):
        e = dict(os.environ); e.update(env)
        p = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ab_time.py"), "--child"],
                           capture_output=True, text=True, env=e)
        res = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        print(sorted(env) or "new rule", res[0][7:] if res else p.stderr[-300:], flush=True)
