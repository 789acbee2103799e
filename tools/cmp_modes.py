"""rel-L2 between chain variants after a few steps (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import __graft_entry__ as ge
from tools.quick_perf import spectrum
nsb = ge.load_package()
n = int(sys.argv[1]); steps = int(sys.argv[2])
HV = dict(enabled=1, coefficient=5.0, order=4, cutoff_ratio=0.3)
res = {}
for name, env in [("plain", dict(NSB_CHAIN="1")), ("plain_oldB", dict(NSB_CHAIN="1", NSB_CHAIN_B512="0")),
                  ("plain_oldG", dict(NSB_CHAIN="1", NSB_CHAIN_G2="0")), ("persistent", dict(NSB_CHAIN="2")),
                  ("generic", dict(NSB_DISABLE_FAST="1"))]:
    for k in ("NSB_CHAIN", "NSB_CHAIN_B512", "NSB_CHAIN_G2", "NSB_DISABLE_FAST"):
        os.environ.pop(k, None)
    os.environ.update(env)
    c = nsb.NsbContext(n, n, Re=1e4)
    c.upload(nsb.OMEGA_HAT, spectrum(n, kk=96))
    c.set_hyperviscosity(**HV)
    for _ in range(steps):
        c.step(0.0)
    res[name] = (c.download(nsb.OMEGA_HAT), c.state()[0])
    c.close()
ref = res["generic"][0]
for k, (w, dt) in res.items():
    print(f"{k:12s} vs generic {np.linalg.norm(w - ref) / np.linalg.norm(ref):.3e}  vs plain "
          f"{np.linalg.norm(w - res['plain'][0]) / np.linalg.norm(ref):.3e}  dt {dt!r} max|diff| {np.abs(w - ref).max():.3e}")
