"""Run a few RK4 steps at one size (target for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
from tools.quick_perf import spectrum
nsb = ge.load_package()
n = int(sys.argv[1]); steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
c = nsb.NsbContext(n, n, Re=1e4)
c.upload(nsb.OMEGA_HAT, spectrum(n))
c.set_hyperviscosity(1, 5.0, 4, 0.3)
for _ in range(steps):
    c.step(0.0)
c.sync()
print(c.plan(), c.state())
