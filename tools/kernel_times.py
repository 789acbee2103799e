"""Per-kernel CUDA-event times of the RK4 step at one size (development aid)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
from tools.quick_perf import spectrum

nsb = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lib = nsb.load_nsb(os.environ.get("NSB_LIB_OVERRIDE"))  # dev aid: alternative builds of libnsb.so
c = nsb.NsbContext(n, n, Re=1e4, lib=lib)
c.upload(nsb.OMEGA_HAT, spectrum(n))
c.set_hyperviscosity(1, 5.0, 4, 0.3)
for _ in range(2):
    c.step(0.0)
c.sync()
c.timer_start()
for _ in range(steps):
    c.step(0.0)
total = c.timer_stop() / steps
c.lib.nsb_profile_begin(c.h)
for _ in range(steps):
    c.step(0.0)
ms = (C.c_double * 6)()
cnt = (C.c_int * 6)()
c.lib.nsb_profile_end(c.h, ms, cnt)
names = ["alpha_inv", "beta_inv/B", "gamma/G/chain", "beta_fwd/F", "alpha_fwd(+inv)", "other"]
print(f"n={n} env={ {k: v for k, v in os.environ.items() if k.startswith('NSB_')} } ms/step={total:.3f}")
for k, nm in enumerate(names):
    if cnt[k]:
        print(f"   {nm:18s} launches/step={cnt[k] // steps:3d} ms/launch={ms[k] / cnt[k]:8.3f} ms/step={ms[k] / steps:8.3f}")
c.close()
