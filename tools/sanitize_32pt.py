"""Small grids that run the 32-point-factor kernels (target for compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
nsb = ge.load_package()
for nx, ny, plan in ((32768, 16, dict(x2=1024, y1=4)), (2048, 64, dict(x2=1024, y1=2))):
    rng = np.random.default_rng(1)
    c = nsb.NsbContext(nx, ny, Re=1e3, **plan)
    w = np.zeros((ny, nx // 2 + 1), dtype=np.complex128)
    w[:, :24] = 0.01 * (rng.standard_normal((ny, 24)) + 1j * rng.standard_normal((ny, 24)))
    w[0, 0] = 0
    c.upload(nsb.OMEGA_HAT, w)
    c.apply_dealiasing(nsb.OMEGA_HAT)
    c.set_hyperviscosity(1, 5.0, 4, 0.3)
    for _ in range(2):
        c.step(0.0)
    c.spectral_to_physical([nsb.OMEGA_HAT], [nsb.OMEGA])
    c.physical_to_spectral(nsb.OMEGA, nsb.SCRATCH_HAT)
    c.sync()
    print(nx, ny, c.plan(), c.state(), flush=True)
    c.close()
