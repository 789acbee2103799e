"""Mint the golden vectors under tests/golden/ from oracle/_ref/libns_ref.so, i.e. from the
UNMODIFIED reference sources (/root/reference) built against oracle/shim/fftw3.h.

Run in the build container (needs /root/reference):  python oracle/gen_golden.py
The reference ships no golden vectors of its own, so these are the pins:
  forcing_kat.json   ns_forcing_generate on the reference's own test case
                     (tests/test_solvers.c:239-295: 4x4 grid, k2 = i^2+j^2, band [0.5,2.5],
                     seed 42, target 1.5), white x2 calls and coloured x2 calls, as C99 hex
                     floats, with the RNG state after each call
  traj_tg32.npz      Taylor-Green 32^2, Re=100: omega_hat after steps 1..5 of
                     `dt = cfl(); advance_rk4()` (after the canonical compute_derivatives),
                     dt series, energy, enstrophy
  traj_forced32.npz  random-phase 32^2, Re=750, hyperviscosity (c=0.5, order 4, cutoff 0.3),
                     coloured forcing seed 42: omega_hat after steps 1..4, dt, E, Z, and the
                     forcing arrays used
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def hexlist(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def forcing_kat():
    lib = C.CDLL(orc.lib_path("ref"))
    nkx = nky = 4
    k2 = np.array([[float(i * i + j * j) for i in range(nkx)] for j in range(nky)])
    out = {}
    for kind, t, dts in [("white", 0, [0.1, 0.1]), ("colored", 1, [0.05, 0.07])]:
        cfg = orc.ForcingConfig()
        lib.ns_forcing_config_default(C.byref(cfg))
        cfg.kmin, cfg.kmax, cfg.energy_target, cfg.seed, cfg.type = 0.5, 2.5, 1.5, 42, t
        if t:
            cfg.correlation_time = 0.2
        st = orc.ForcingState()
        lib.ns_forcing_state_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        assert lib.ns_forcing_state_init(C.byref(st), C.byref(cfg), k2.ctypes.data, nkx, nky) == 0
        calls = []
        for dt in dts:
            buf = np.zeros((nky, nkx), dtype=np.complex128)
            lib.ns_forcing_generate.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
            lib.ns_forcing_generate(C.byref(st), dt, k2.ctypes.data, buf.ctypes.data)
            calls.append({"dt": dt, "values": hexlist(buf.view(np.float64)), "rng_state": int(st.rng_state),
                          "have_spare": int(st.have_spare), "energy": float((np.abs(buf) ** 2).sum())})
        out[kind] = {"active_modes": int(st.active_modes), "calls": calls}
    with open(os.path.join(OUT, "forcing_kat.json"), "w") as f:
        json.dump(out, f, indent=1)


def traj_tg():
    o = orc.Oracle("ref", 32, 32, Re=100.0)
    o.taylor_green()
    w, dts, e, z = [o.arr("omega_hat").copy()], [], [o.energy()], [o.enstrophy()]
    for _ in range(5):
        dts.append(o.step())
        w.append(o.arr("omega_hat").copy())
        e.append(o.energy())
        z.append(o.enstrophy())
    np.savez(os.path.join(OUT, "traj_tg32.npz"), omega_hat=np.array(w), dt=np.array(dts), energy=np.array(e),
             enstrophy=np.array(z), t=o.s.t)
    o.close()


def traj_forced():
    o = orc.Oracle("ref", 32, 32, Re=750.0)
    o.set_hyperviscosity(1, 0.5, 4, 0.3)
    o.random_phase(seed=7)
    o.init_forcing("colored", kmin=0.5, kmax=4.0, energy_target=1.25, tau=0.4, seed=42)
    w, dts, e, z, fs = [o.arr("omega_hat").copy()], [], [o.energy()], [o.enstrophy()], []
    for _ in range(4):
        dt = o.compute_cfl_timestep(o.sp)
        f = o.gen_forcing(dt).copy()
        o.step(dt=dt, forcing=f)
        fs.append(f)
        dts.append(dt)
        w.append(o.arr("omega_hat").copy())
        e.append(o.energy())
        z.append(o.enstrophy())
    np.savez(os.path.join(OUT, "traj_forced32.npz"), omega_hat=np.array(w), dt=np.array(dts),
             energy=np.array(e), enstrophy=np.array(z), forcing=np.array(fs),
             rng_state=np.uint64(o.forcing_state.rng_state))
    o.close()


if __name__ == "__main__":
    assert orc.have("ref"), "build oracle/_ref first (make -C oracle ref; needs /root/reference)"
    os.makedirs(OUT, exist_ok=True)
    forcing_kat()
    traj_tg()
    traj_forced()
    print("golden vectors written to", OUT)
