"""Quick device-side timing of the RK4 step at several sizes (development aid)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

nsb = ge.load_package()


def spectrum(n, seed=0, kk=48):
    rng = np.random.default_rng(seed)
    nkx = n // 2 + 1
    w = np.zeros((n, nkx), dtype=np.complex128)
    blk = rng.standard_normal((2 * kk + 1, kk + 1)) + 1j * rng.standard_normal((2 * kk + 1, kk + 1))
    rows = np.arange(-kk, kk + 1) % n
    w[rows[:, None], np.arange(kk + 1)[None, :]] = 0.02 * blk
    w[0, 0] = 0
    col = w[:, 0].copy()
    w[:, 0] = 0.5 * (col + np.conj(col[(-np.arange(n)) % n]))
    return w


def run(n, steps, **plan):
    c = nsb.NsbContext(n, n, Re=1e4, **plan)
    c.upload(nsb.OMEGA_HAT, spectrum(n))
    c.set_hyperviscosity(1, 5.0, 4, 0.3)
    for _ in range(2):
        c.step(0.0)
    c.sync()
    c.timer_start()
    for _ in range(steps):
        c.step(0.0)
    ms = c.timer_stop() / steps
    gps = n * n / (ms * 1e-3)
    print(f"n={n} plan={c.plan()} ms/step={ms:.3f} gp-steps/s={gps:.3e} "
          f"model GB/s={512 * gps / 1e9:.0f} frac(6552)={512 * gps / 6552.3e9:.3f} dt={c.state()[0]:.3e}", flush=True)
    c.close()


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [1024, 4096, 16384]
    for n in sizes:
        run(n, 20 if n <= 4096 else 5)
