"""Shared bodies of the parity tests: the same checks run against the CUDA library
(-m gpu) and against the kernel sources under the CPU emulator (-m "not gpu")."""
import numpy as np

TOL_STEP = 1e-11      # north-star: rel L2 on omega_hat per step
TOL_DIAG = 1e-9       # north-star: energy / enstrophy relative


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def check_transforms(nsb, lib, orc, kind, nx, ny, **plan):
    c = nsb.NsbContext(nx, ny, lib=lib, **plan)
    o = orc.Oracle(kind, nx, ny)
    rng = np.random.default_rng(1)
    f = rng.standard_normal((ny, nx))
    c.upload(nsb.OMEGA, f)
    c.physical_to_spectral(nsb.OMEGA, nsb.SCRATCH_HAT)
    o.arr("omega")[...] = f
    o.physical_to_spectral(o.sp, o.s.omega, o.s.omega_hat)
    assert rel(c.download(nsb.SCRATCH_HAT), o.arr("omega_hat")) < 1e-14
    # general spectrum incl. non-Hermitian DC column and non-zero masked modes
    g = rng.standard_normal((ny, nx // 2 + 1)) + 1j * rng.standard_normal((ny, nx // 2 + 1))
    c.upload(nsb.SCRATCH_HAT, g)
    c.spectral_to_physical([nsb.SCRATCH_HAT], [nsb.PSI])
    o.arr("omega_hat")[...] = g
    o.spectral_to_physical(o.sp, o.s.omega_hat, o.s.omega)
    assert rel(c.download(nsb.PSI), o.arr("omega")) < 1e-14
    c.close()
    o.close()


def make_pair(nsb, lib, orc, kind, nx, ny, Re=750.0, hv=None, forcing=None, init="random", **plan):
    o = orc.Oracle(kind, nx, ny, Re=Re)
    c = nsb.NsbContext(nx, ny, Re=Re, lib=lib, **plan)
    if init == "random":
        o.random_phase()
    else:
        o.taylor_green()
    if hv:
        o.set_hyperviscosity(**hv)
    if forcing:
        o.init_forcing(**forcing)
    c.upload(nsb.OMEGA_HAT, o.arr("omega_hat"))
    if hv:
        c.set_hyperviscosity(**hv)
    return o, c


def check_operators(nsb, lib, orc, kind, nx, ny, **plan):
    """velocity-from-vorticity, compute_derivatives, nonlinear terms vs the oracle."""
    o, c = make_pair(nsb, lib, orc, kind, nx, ny, **plan)
    c.upload(nsb.PSI_HAT, o.arr("psi_hat"))
    o.compute_nonlinear_terms(o.sp)
    c.nonlinear_terms()
    assert rel(c.download(nsb.NONLINEAR_HAT), o.arr("nonlinear_hat")) < 1e-13
    o.compute_derivatives(o.sp)
    c.compute_derivatives()
    for name, fid in [("psi_hat", nsb.PSI_HAT), ("u_hat", nsb.U_HAT), ("v_hat", nsb.V_HAT)]:
        assert rel(c.download(fid), o.arr(name)) < 1e-15, name
    for name, fid in [("u", nsb.U), ("v", nsb.V), ("omega", nsb.OMEGA)]:
        assert rel(c.download(fid), o.arr(name)) < 1e-13, name
    d = c.diagnostics()
    assert abs(d["energy"] - o.energy()) <= 1e-13 * o.energy()
    assert abs(d["enstrophy"] - o.enstrophy()) <= 1e-13 * o.enstrophy()
    assert abs(c.cfl_timestep() - o.compute_cfl_timestep(o.sp)) <= 1e-14
    c.close()
    o.close()


def check_per_step(nsb, lib, orc, kind, nx, ny, nsteps, hv=None, forcing=None, device_dt=True,
                   init="random", Re=750.0, **plan):
    """Per-step parity: both sides start every step from the oracle's omega_hat."""
    o, c = make_pair(nsb, lib, orc, kind, nx, ny, Re=Re, hv=hv, forcing=forcing, init=init, **plan)
    worst = 0.0
    for _ in range(nsteps):
        c.upload(nsb.OMEGA_HAT, o.arr("omega_hat"))
        if hv and hv.get("sfd_enabled"):
            c.upload(nsb.SFD_HAT, o.view.cplx("sfd_hat"))
        dt = o.compute_cfl_timestep(o.sp)
        f = o.gen_forcing(dt) if forcing else None
        if forcing:
            c.set_forcing_dense(f)
        o.step(dt=dt, forcing=f)
        c.step(0.0 if device_dt else dt)
        err = rel(c.download(nsb.OMEGA_HAT), o.arr("omega_hat"))
        worst = max(worst, err)
        assert err <= TOL_STEP, err
        if device_dt:
            assert abs(c.state()[0] - dt) <= 1e-13 * dt
    if hv and hv.get("sfd_enabled"):
        assert rel(c.download(nsb.SFD_HAT), o.view.cplx("sfd_hat")) <= TOL_STEP
    c.close()
    o.close()
    return worst


def check_free_running(nsb, lib, orc, kind, nx, ny, nsteps, every=10, hv=None, forcing=None,
                       init="random", Re=750.0, **plan):
    """Free-running: each side uses its own adaptive dt; E and enstrophy must track."""
    o, c = make_pair(nsb, lib, orc, kind, nx, ny, Re=Re, hv=hv, forcing=forcing, init=init, **plan)
    worst = 0.0
    for n in range(1, nsteps + 1):
        if forcing:
            dt_gpu = c.cfl_timestep() if forcing.get("kind") == "colored" else 0.0
            # the stream is consumed identically on both sides: oracle generates, GPU replays
            dt = o.compute_cfl_timestep(o.sp)
            f = o.gen_forcing(dt)
            c.set_forcing_dense(f)
            o.step(dt=dt, forcing=f)
            c.step(dt_gpu)
        else:
            o.step()
            c.step(0.0)
        if n % every == 0:
            d = c.diagnostics()
            e, z = o.energy(), o.enstrophy()
            worst = max(worst, abs(d["energy"] - e) / e, abs(d["enstrophy"] - z) / z)
            assert abs(d["energy"] - e) <= TOL_DIAG * e, (n, d["energy"], e)
            assert abs(d["enstrophy"] - z) <= TOL_DIAG * z, (n, d["enstrophy"], z)
    assert abs(c.state()[1] - o.s.t) <= 1e-9 * o.s.t
    c.close()
    o.close()
    return worst


def check_spectral_stats(nsb, lib, nx, ny, **plan):
    """nsb_spectral_stats against numpy over the same omega_hat (the reductions of
    spectral_analysis.c:36-66 and :69-100)."""
    rng = np.random.default_rng(17)
    c = nsb.NsbContext(nx, ny, lib=lib, **plan)
    w = rng.standard_normal((ny, nx // 2 + 1)) + 1j * rng.standard_normal((ny, nx // 2 + 1))
    c.upload(nsb.OMEGA_HAT, w)
    st = c.spectral_stats()
    k2 = c.k2()
    k2_last = k2.reshape(-1)[-1]
    mag = np.abs(w)
    kn = np.sqrt(k2) / np.sqrt(k2_last)
    assert abs(st["k2_last"] - k2_last) <= 1e-12 * k2_last
    assert abs(st["total"] - (mag ** 2).sum()) <= 1e-12 * (mag ** 2).sum()
    assert abs(st["high_k"] - (mag[kn > 0.8] ** 2).sum()) <= 1e-12 * (mag ** 2).sum()
    assert abs(st["max_abs"] - mag.max()) <= 1e-14 * mag.max()
    assert abs(st["tail"] - (mag[k2 > 0.25 * k2_last] ** 2).sum()) <= 1e-12 * (mag ** 2).sum()
    assert st["tail_modes"] == float((k2 > 0.25 * k2_last).sum())
    # nsb_energy_spectrum against numpy (spectral_analysis.c:103-142 over spectral_core.c:8-34)
    nb = min(nx // 2 + 1, ny // 2)
    ek, cnt = c.energy_spectrum(nb)
    kx = (2.0 * np.pi * np.arange(nx // 2 + 1) / c.Lx)[None, :]
    jj = np.arange(ny)
    ky = np.where(jj <= ny // 2, 2.0 * np.pi * jj / c.Ly, 2.0 * np.pi * (jj - ny) / c.Ly)[:, None]
    psi = np.where(k2 > 1e-12, -w / np.where(k2 > 1e-12, k2, 1.0), 0.0)
    e = 0.5 * (np.abs(ky * psi) ** 2 + np.abs(kx * psi) ** 2)
    bins = (np.sqrt(k2) + 0.5).astype(np.int64)
    ok = bins < nb
    assert np.array_equal(cnt, np.bincount(bins[ok], minlength=nb))
    want = np.bincount(bins[ok], weights=e[ok], minlength=nb)
    assert np.max(np.abs(ek - want)) <= 1e-12 * want.max()
    c.close()


def check_parseval(nsb, lib, nx, ny, **plan):
    """nsb_diagnostics_spectral (Parseval over the retained modes, Hermitian projection of the kx = 0 column) against
    nsb_diagnostics (three c2r transforms + sums over the grid, spectral_analysis.c:8-33) on a GENERAL omega_hat:
    non-Hermitian kx = 0 column, non-zero masked modes."""
    rng = np.random.default_rng(23)
    c = nsb.NsbContext(nx, ny, lib=lib, **plan)
    w = rng.standard_normal((ny, nx // 2 + 1)) + 1j * rng.standard_normal((ny, nx // 2 + 1))
    c.upload(nsb.OMEGA_HAT, w)
    a, b = c.diagnostics(), c.diagnostics_spectral()
    assert abs(a["energy"] - b["energy"]) <= 1e-13 * a["energy"], (a, b)
    assert abs(a["enstrophy"] - b["enstrophy"]) <= 1e-13 * a["enstrophy"], (a, b)
    c.close()
