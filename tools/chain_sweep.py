"""Sweep the chain mode / ring parameters at one size (development aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_perf import run

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
configs = [dict(NSB_CHAIN="1")]
for ctas in (2, 1):
    for z, lg, p, lf in [(6, 2, 10, 6), (10, 2, 16, 10), (16, 3, 24, 12), (24, 4, 32, 16)]:
        configs.append(dict(NSB_CHAIN="2", NSB_CHAIN_CTAS=str(ctas), NSB_CHAIN_SLOTS_Z=str(z), NSB_CHAIN_LEAD_G=str(lg),
                            NSB_CHAIN_SLOTS_P=str(p), NSB_CHAIN_LEAD_F=str(lf)))
for cfg in configs:
    for k in list(os.environ):
        if k.startswith("NSB_CHAIN"):
            del os.environ[k]
    os.environ.update(cfg)
    print(cfg, flush=True)
    run(n, steps)
