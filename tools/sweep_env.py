"""ms/step at one size for a list of NSB_* environment settings (development aid).
usage: sweep_env.py n steps "A=1 B=2" "A=3" ..."""
import os
import subprocess
import sys

n, steps = sys.argv[1], sys.argv[2]
here = os.path.dirname(os.path.abspath(__file__))
for spec in sys.argv[3:]:
    env = dict(os.environ)
    for kv in spec.split():
        k, v = kv.split("=")
        env[k] = v
    out = subprocess.run([sys.executable, os.path.join(here, "quick_perf.py"), n], env=env, capture_output=True, text=True)
    line = [l for l in out.stdout.split("\n") if "ms/step" in l]
    print(f"{spec:60s} {line[0].split('ms/step=')[1].split()[0] if line else out.stderr[-300:]}", flush=True)
