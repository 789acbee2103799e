"""The reference's own passing tests (tests/test_solvers.c, tests/test_conservation.c),
re-stated against the drop-in host library. Run on the CPU through the emulator-linked
build (not gpu) and on the GPU through the product build (gpu)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

import hostlib
import orc
from conftest import ROOT, _clean_env, best_oracle_kind

TWO_PI = 2 * math.pi


@pytest.fixture(scope="module")
def emu_host(emu_lib):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu"), "all"], check=True, env=_clean_env(),
                   stdout=subprocess.DEVNULL)
    return hostlib.load(hostlib.EMU)


@pytest.fixture(scope="module")
def gpu_host(cuda_lib):
    return hostlib.load(hostlib.PRODUCT)


def _reference_unit_tests(lib):
    # test_spectral_allocation (test_solvers.c:433-448)
    s = hostlib.Solver(lib, 64, 64, plans=False)
    assert s.s.nx == 64 and s.s.nkx == 33 and s.s.u and s.s.omega_hat
    s.close()
    # test_spectral_grid_initialization (:450-471)
    s = hostlib.Solver(lib, 32, 32, plans=False)
    kx = s.view.vec("kx", 17)
    assert abs(kx[0]) < 1e-10 and abs(kx[1] - 1.0) < 1e-10
    assert abs(s.view.k2[0, 0]) < 1e-10 and s.view.mask[0, 0] == 1
    assert int((s.view.mask == 0).sum()) == 292          # 292 of 544 masked at 32^2
    s.close()
    # test_spectral_fft_transforms (:473-511)
    s = hostlib.Solver(lib, 32, 32)
    x = s.view.vec("x", 32)
    y = s.view.vec("y", 32)
    s.arr("omega")[...] = np.sin(x)[None, :] * np.cos(y)[:, None]
    orig = s.arr("omega").copy()
    lib.ns_spectral_physical_to_spectral(s.sp, s.s.omega, s.s.omega_hat)
    lib.ns_spectral_spectral_to_physical(s.sp, s.s.omega_hat, s.s.omega)
    assert np.abs(s.arr("omega") - orig).max() < 1e-10
    # test_spectral_dealiasing (:513-545)
    s.arr("omega_hat")[...] = 1.0
    lib.ns_spectral_apply_dealiasing(s.sp, s.s.omega_hat)
    w = s.arr("omega_hat")
    assert np.all(np.abs(w[s.view.mask == 0]) < 1e-15) and np.all(w[s.view.mask == 1] == 1.0)
    # test_spectral_velocity_from_vorticity (:547-577)
    s.arr("omega_hat")[...] = 0.0
    s.arr("omega_hat")[0, 0] = 1.0
    lib.ns_spectral_compute_velocity_from_vorticity(s.sp)
    lib.ns_spectral_spectral_to_physical(s.sp, s.s.u_hat, s.s.u)
    lib.ns_spectral_spectral_to_physical(s.sp, s.s.v_hat, s.s.v)
    assert max(np.abs(s.arr("u")).max(), np.abs(s.arr("v")).max()) < 1e-10
    # test_spectral_energy/enstrophy_computation (:579-614); values from SURVEY section 4
    lib.ns_spectral_initialize_taylor_green(s.sp)
    e, z = lib.ns_spectral_compute_energy(s.sp), lib.ns_spectral_compute_enstrophy(s.sp)
    assert 0 < e < 1.0 and abs(e - 0.25) < 1e-12 and abs(z - 0.5) < 1e-12
    s.close()
    # test_spectral_hyperviscosity (:616-664)
    s = hostlib.Solver(lib, 32, 32)
    h = orc.Hyperviscosity()
    lib.ns_spectral_hyperviscosity_default(C.byref(h))
    h.enabled, h.coefficient, h.order, h.cutoff_ratio = 1, 0.05, 4, 0.6
    h.sfd_enabled, h.sfd_strength, h.sfd_timescale = 1, 0.3, 0.4
    lib.ns_spectral_hyperviscosity_set(s.sp, C.byref(h))
    assert abs(s.view.weights[0, 0]) < 1e-12 and s.view.weights[-1, -1] > 0.0
    s.arr("rk1_hat")[...] = 0.0
    s.arr("omega_hat")[...] = 1.0
    lib.ns_spectral_hyperviscosity_apply(s.sp, s.s.rk1_hat, s.s.omega_hat)
    rhs = s.arr("rk1_hat")
    assert abs(rhs[0, 0]) < 1e-8 and abs(rhs[-1, -1]) > 1e-3
    lib.ns_spectral_hyperviscosity_post_step(s.sp, s.s.omega_hat, 0.1)
    assert abs(s.view.cplx("sfd_hat")[-1, -1]) > 0.0
    s.close()


def _conservation_tests(lib, orc_mod, steps=30):
    """test_spectral_incompressibility + energy/enstrophy decay (test_conservation.c:62-186),
    and trajectory parity of the EAGER entry points with the oracle."""
    nx = ny = 64
    s = hostlib.Solver(lib, nx, ny, Re=100.0)
    o = orc_mod.Oracle(best_oracle_kind(orc_mod), nx, ny, Re=100.0)
    lib.ns_spectral_initialize_taylor_green(s.sp)
    o.initialize_taylor_green(o.sp)
    lib.ns_spectral_compute_derivatives(s.sp)   # canonical start
    o.compute_derivatives(o.sp)
    e0, z0 = lib.ns_spectral_compute_energy(s.sp), lib.ns_spectral_compute_enstrophy(s.sp)
    for _ in range(steps):
        s.s.dt = lib.ns_spectral_compute_cfl_timestep(s.sp)
        lib.ns_spectral_advance_rk4(s.sp)
        o.step()
        u, v = s.arr("u"), s.arr("v")
        du = (np.roll(u, -1, axis=1) - np.roll(u, 1, axis=1)) / (2 * TWO_PI / nx)
        dv = (np.roll(v, -1, axis=0) - np.roll(v, 1, axis=0)) / (2 * TWO_PI / ny)
        assert np.abs(du + dv).max() < 1e-12
        assert abs(s.s.dt - o.s.dt) <= 1e-15
    e1, z1 = lib.ns_spectral_compute_energy(s.sp), lib.ns_spectral_compute_enstrophy(s.sp)
    assert 0 < e1 < e0 and 0 < z1 < z0
    w, wo = s.arr("omega_hat"), o.arr("omega_hat")
    assert np.linalg.norm(w - wo) / np.linalg.norm(wo) < 1e-13
    assert abs(e1 - o.energy()) < 1e-13 and abs(s.s.t - o.s.t) < 1e-14
    for name in ["psi_hat", "u_hat", "v_hat", "u", "v", "omega"]:
        a, b = s.arr(name), o.arr(name)
        assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-12, name
    s.close()
    o.close()


def _lazy_run_steps(lib, orc_mod, kind):
    """ns_spectral_run_steps (device-resident loop, forcing replayed from the seeded
    stream) against the oracle loop, white and coloured."""
    nx, ny = 64, 32
    s = hostlib.Solver(lib, nx, ny, Re=750.0)
    o = orc_mod.Oracle(best_oracle_kind(orc_mod), nx, ny, Re=750.0)
    o.random_phase(seed=11)
    lib.ns_spectral_initialize_random_phase(s.sp, 11, 8.0, 5.0)
    assert np.linalg.norm(s.arr("omega_hat") - o.arr("omega_hat")) / np.linalg.norm(o.arr("omega_hat")) < 1e-14
    s.arr("omega_hat")[...] = o.arr("omega_hat")
    hv = dict(enabled=1, coefficient=0.5, order=4, cutoff_ratio=0.3)
    o.set_hyperviscosity(**hv)
    h = orc.Hyperviscosity()
    lib.ns_spectral_hyperviscosity_default(C.byref(h))
    h.enabled, h.coefficient, h.order, h.cutoff_ratio = 1, 0.5, 4, 0.3
    lib.ns_spectral_hyperviscosity_set(s.sp, C.byref(h))
    o.init_forcing(kind)
    cfg = orc.ForcingConfig()
    lib.ns_forcing_config_default(C.byref(cfg))
    cfg.kmin, cfg.kmax, cfg.energy_target, cfg.correlation_time, cfg.seed = 0.5, 4.0, 1.25, 0.4, 42
    cfg.type = 1 if kind == "colored" else 0
    fs = orc.ForcingState()
    assert lib.ns_forcing_state_init(C.byref(fs), C.byref(cfg), s.s.k2, s.s.nkx, s.s.nky) == 0
    lib.ns_spectral_set_sync_policy(s.sp, 1)
    assert lib.ns_spectral_run_steps(s.sp, 6, C.byref(fs)) == 0
    for _ in range(6):
        o.step()
    lib.ns_spectral_sync_to_host(s.sp, 0xFFFF)
    assert fs.rng_state == o.forcing_state.rng_state      # stream consumed identically
    err = np.linalg.norm(s.arr("omega_hat") - o.arr("omega_hat")) / np.linalg.norm(o.arr("omega_hat"))
    assert err < 1e-12, err
    assert abs(s.s.t - o.s.t) < 1e-13 and abs(s.s.dt - o.s.dt) < 1e-15
    assert abs(lib.ns_spectral_compute_energy(s.sp) - o.energy()) < 1e-13
    lib.ns_forcing_state_free(C.byref(fs))
    s.close()
    o.close()


def _forced_solver(lib, kind, nx=64, ny=32, sfd=False):
    """solver + forcing state of the forced test case (random-phase IC, hyperviscosity, band [0.5, 4])"""
    s = hostlib.Solver(lib, nx, ny, Re=750.0)
    lib.ns_spectral_initialize_random_phase(s.sp, 11, 8.0, 5.0)
    h = orc.Hyperviscosity()
    lib.ns_spectral_hyperviscosity_default(C.byref(h))
    h.enabled, h.coefficient, h.order, h.cutoff_ratio = 1, 0.5, 4, 0.3
    if sfd:
        h.sfd_enabled, h.sfd_strength, h.sfd_timescale = 1, 0.2, 0.05
    lib.ns_spectral_hyperviscosity_set(s.sp, C.byref(h))
    cfg = orc.ForcingConfig()
    lib.ns_forcing_config_default(C.byref(cfg))
    cfg.kmin, cfg.kmax, cfg.energy_target, cfg.correlation_time, cfg.seed = 0.5, 4.0, 1.25, 0.4, 42
    cfg.type = 1 if kind == "colored" else 0
    fs = orc.ForcingState()
    assert lib.ns_forcing_state_init(C.byref(fs), C.byref(cfg), s.s.k2, s.s.nkx, s.s.nky) == 0
    return s, fs


def _restart_bit_exact(lib, tmp_path, kind):
    """4 steps, save, 5 more == load into a fresh solver, 5 steps, bit for bit
    (omega_hat, SFD state, t, dt, forcing stream position, coloured-noise state)."""
    path = str(tmp_path / ("restart_%s.bin" % kind)).encode()
    a, fa = _forced_solver(lib, kind, sfd=True)
    lib.ns_spectral_set_sync_policy(a.sp, 1)
    assert lib.ns_spectral_run_steps(a.sp, 4, C.byref(fa)) == 0
    assert lib.ns_spectral_save_state(a.sp, C.byref(fa), path) == 0, lib.ns_spectral_last_error()
    assert lib.ns_spectral_run_steps(a.sp, 5, C.byref(fa)) == 0
    lib.ns_spectral_sync_to_host(a.sp, 0xFFFF)

    b, fb = _forced_solver(lib, kind, sfd=True)
    b.arr("omega_hat")[...] = 0.0          # everything must come from the file
    lib.ns_spectral_set_sync_policy(b.sp, 1)
    assert lib.ns_spectral_load_state(b.sp, C.byref(fb), path) == 0, lib.ns_spectral_last_error()
    assert lib.ns_spectral_run_steps(b.sp, 5, C.byref(fb)) == 0
    lib.ns_spectral_sync_to_host(b.sp, 0xFFFF)
    assert fb.rng_state == fa.rng_state and fb.have_spare == fa.have_spare
    assert b.s.t == a.s.t and b.s.dt == a.s.dt
    assert np.array_equal(b.arr("omega_hat"), a.arr("omega_hat"))
    assert np.array_equal(b.arr("sfd_hat"), a.arr("sfd_hat"))
    assert np.linalg.norm(a.arr("sfd_hat")) > 0

    # error paths: wrong grid, not a restart file
    c = hostlib.Solver(lib, 32, 32, Re=750.0)
    assert lib.ns_spectral_load_state(c.sp, None, path) == -1
    assert b"grid" in lib.ns_spectral_last_error()
    junk = tmp_path / "junk.bin"
    junk.write_bytes(b"not a restart file" * 10)
    assert lib.ns_spectral_load_state(c.sp, None, str(junk).encode()) == -1
    c.close()
    for s_, f_ in ((a, fa), (b, fb)):
        lib.ns_forcing_state_free(C.byref(f_))
        s_.close()


def _spectrum_and_strided_output(lib, tmp_path):
    """device shell spectrum (LAZY) == host loop of spectral_analysis.c:103-142
    (EAGER) == numpy; strided dump = every stride-th line of the reference-format dump."""
    s, fs = _forced_solver(lib, "white", nx=64, ny=64)
    lib.ns_spectral_compute_derivatives(s.sp)
    nb = min(s.s.nkx, s.s.nky // 2)
    ek_h, cnt_h = np.zeros(nb), np.zeros(nb, dtype=np.int32)
    ek_d, cnt_d = np.zeros(nb), np.zeros(nb, dtype=np.int32)
    ip = C.POINTER(C.c_int)
    assert lib.ns_spectral_energy_spectrum(s.sp, nb, ek_h.ctypes.data_as(hostlib.c_double_p),
                                           cnt_h.ctypes.data_as(ip)) == 0
    lib.ns_spectral_set_sync_policy(s.sp, 1)
    assert lib.ns_spectral_energy_spectrum(s.sp, nb, ek_d.ctypes.data_as(hostlib.c_double_p),
                                           cnt_d.ctypes.data_as(ip)) == 0
    k2 = np.ctypeslib.as_array(s.s.k2, shape=(s.s.nky * s.s.nkx,))
    bins = (np.sqrt(k2) + 0.5).astype(np.int64)
    e = 0.5 * (np.abs(s.arr("u_hat").ravel()) ** 2 + np.abs(s.arr("v_hat").ravel()) ** 2)
    ok = bins < nb
    ek_n = np.bincount(bins[ok], weights=e[ok], minlength=nb)
    cnt_n = np.bincount(bins[ok], minlength=nb)
    assert np.array_equal(cnt_h, cnt_n) and np.array_equal(cnt_d, cnt_n)
    assert ek_n.max() > 0
    assert np.max(np.abs(ek_h - ek_n)) <= 1e-13 * ek_n.max()
    assert np.max(np.abs(ek_d - ek_n)) <= 1e-13 * ek_n.max()

    full, thin = tmp_path / "full.dat", tmp_path / "thin.dat"
    lib.ns_spectral_output_solution(s.sp, str(full).encode())
    assert lib.ns_spectral_output_solution_strided(s.sp, str(thin).encode(), 4) == 0
    assert lib.ns_spectral_output_solution_strided(s.sp, str(thin).encode(), 0) == -1
    fl = full.read_text().split("\n")
    assert fl[0] == "# Spectral Navier-Stokes Solution" and fl[1] == "# x y u v omega psi"
    nx, ny = s.s.nx, s.s.ny
    want = fl[:2]
    for j in range(0, ny, 4):
        row0 = 2 + j * (nx + 1)
        want += fl[row0:row0 + nx:4] + [""]
    assert thin.read_text().split("\n") == want + [""]
    lib.ns_forcing_state_free(C.byref(fs))
    s.close()


def _blowup_is_reported(lib, n):
    """a NaN in the field is carried through the fused max|u| reduction to the CFL kernel and
    ns_spectral_run_steps fails loudly instead of integrating on with dt = MAX_DT like the reference would."""
    s = hostlib.Solver(lib, n, n, Re=500.0)
    lib.ns_spectral_initialize_random_phase(s.sp, 3, 6.0, 3.0)
    lib.ns_spectral_set_sync_policy(s.sp, 1)
    assert lib.ns_spectral_run_steps(s.sp, 2, None) == 0
    lib.ns_spectral_sync_to_host(s.sp, 0x1)                      # NS_FIELD_OMEGA_HAT
    s.arr("omega_hat")[2, 1] = complex(np.nan, 0.0)
    lib.ns_spectral_sync_to_device(s.sp, 0x1)
    assert lib.ns_spectral_run_steps(s.sp, 3, None) == -1
    assert b"no longer finite" in lib.ns_spectral_last_error()
    # a fresh, finite state clears the verdict
    lib.ns_spectral_initialize_random_phase(s.sp, 3, 6.0, 3.0)
    assert lib.ns_spectral_run_steps(s.sp, 1, None) == 0
    s.close()


def test_blowup_is_reported_emu(emu_host):
    _blowup_is_reported(emu_host, 32)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [64, 2048])
def test_blowup_is_reported_gpu(gpu_host, n):
    _blowup_is_reported(gpu_host, n)


def test_reference_unit_tests_emu(emu_host):
    _reference_unit_tests(emu_host)


def test_conservation_and_eager_parity_emu(emu_host, oracles):
    _conservation_tests(emu_host, oracles, steps=6)


@pytest.mark.parametrize("kind", ["white", "colored"])
def test_lazy_run_steps_emu(emu_host, oracles, kind):
    _lazy_run_steps(emu_host, oracles, kind)


@pytest.mark.parametrize("kind", ["white", "colored"])
def test_restart_bit_exact_emu(emu_host, tmp_path, kind):
    _restart_bit_exact(emu_host, tmp_path, kind)


def test_spectrum_and_strided_output_emu(emu_host, tmp_path):
    _spectrum_and_strided_output(emu_host, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["white", "colored"])
def test_restart_bit_exact_gpu(gpu_host, tmp_path, kind):
    _restart_bit_exact(gpu_host, tmp_path, kind)


@pytest.mark.gpu
def test_spectrum_and_strided_output_gpu(gpu_host, tmp_path):
    _spectrum_and_strided_output(gpu_host, tmp_path)


@pytest.mark.gpu
def test_reference_unit_tests_gpu(gpu_host):
    _reference_unit_tests(gpu_host)


@pytest.mark.gpu
def test_conservation_and_eager_parity_gpu(gpu_host, oracles):
    _conservation_tests(gpu_host, oracles, steps=30)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["white", "colored"])
def test_lazy_run_steps_gpu(gpu_host, oracles, kind):
    _lazy_run_steps(gpu_host, oracles, kind)


def _group_case(lib, n, gpus, case):
    """one scenario of the reference-facing entry points, on one device or on a device group"""
    s = hostlib.Solver(lib, n, n, Re=1e4)
    if gpus > 1:
        assert lib.ns_spectral_backend_init(s.sp, gpus) == 0, lib.ns_spectral_last_error()
    lib.ns_spectral_initialize_random_phase(s.sp, 11, 8.0, 5.0)
    h = orc.Hyperviscosity()
    lib.ns_spectral_hyperviscosity_default(C.byref(h))
    h.enabled, h.coefficient, h.order, h.cutoff_ratio = 1, 5.0, 4, 0.3
    lib.ns_spectral_hyperviscosity_set(s.sp, C.byref(h))
    out = {}
    if case == "eager_loop":
        # the reference driver loop (main_spectral.c:46-48) under NS_SYNC_EAGER
        for _ in range(2):
            s.s.dt = lib.ns_spectral_compute_cfl_timestep(s.sp)
            lib.ns_spectral_advance_rk4(s.sp)
        out["dt"], out["t"] = s.s.dt, s.s.t
    elif case == "lazy_sfd_after_steps":
        lib.ns_spectral_set_sync_policy(s.sp, 1)
        assert lib.ns_spectral_run_steps(s.sp, 2, None) == 0
        h.sfd_enabled, h.sfd_strength, h.sfd_timescale = 1, 0.3, 0.05
        lib.ns_spectral_hyperviscosity_set(s.sp, C.byref(h))      # seeds sfd_hat from the CURRENT omega_hat
        assert lib.ns_spectral_run_steps(s.sp, 2, None) == 0
        # reductions over omega_hat right after group steps: partial sums per device, no gather
        nb = 64
        ek, cnt = np.zeros(nb), np.zeros(nb, dtype=np.int32)
        assert lib.ns_spectral_energy_spectrum(s.sp, nb, ek.ctypes.data_as(hostlib.c_double_p),
                                               cnt.ctypes.data_as(C.POINTER(C.c_int))) == 0
        out["approx_ek"], out["cnt"] = ek, cnt
        out["approx_energy"] = np.array([lib.ns_spectral_compute_energy(s.sp), lib.ns_spectral_compute_enstrophy(s.sp)])
        out["resolved"] = lib.ns_spectral_check_resolution(s.sp)
        out["cfl"] = lib.ns_spectral_compute_cfl_timestep(s.sp)   # gathered single-device call on a group
        lib.ns_spectral_sync_to_host(s.sp, 0xFFFFFFFF)
        out["dt"], out["t"] = s.s.dt, s.s.t
        out["sfd"] = np.array(s.arr("sfd_hat"))
        assert np.linalg.norm(out["sfd"]) > 0
    else:  # coloured forcing: the CFL step is needed on the host before every step
        cfg = orc.ForcingConfig()
        lib.ns_forcing_config_default(C.byref(cfg))
        cfg.kmin, cfg.kmax, cfg.energy_target, cfg.correlation_time, cfg.seed, cfg.type = 0.5, 4.0, 1.25, 0.4, 42, 1
        fs = orc.ForcingState()
        assert lib.ns_forcing_state_init(C.byref(fs), C.byref(cfg), s.s.k2, s.s.nkx, s.s.nky) == 0
        lib.ns_spectral_set_sync_policy(s.sp, 1)
        assert lib.ns_spectral_run_steps(s.sp, 3, C.byref(fs)) == 0, lib.ns_spectral_last_error()
        lib.ns_spectral_sync_to_host(s.sp, 0xFFFFFFFF)
        out["dt"], out["t"], out["rng"] = s.s.dt, s.s.t, fs.rng_state
        lib.ns_forcing_state_free(C.byref(fs))
    out["w"] = np.array(s.arr("omega_hat"))
    s.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["eager_loop", "lazy_sfd_after_steps", "colored_forcing"])
def test_device_group_entry_points_gpu(gpu_host, monkeypatch, case):
    """Every entry point on a slab-decomposed device group (ns_spectral_backend_init, here 2 logical devices
    on GPU 0) must reproduce the single-device run bit for bit: the reference driver loop cfl + advance_rk4
    (EAGER), enabling the SFD filter after LAZY steps, and coloured forcing through ns_spectral_run_steps."""
    monkeypatch.setenv("NS_SPECTRAL_GROUP_SAME_DEVICE", "1")
    one = _group_case(gpu_host, 2048, 1, case)
    grp = _group_case(gpu_host, 2048, 2, case)
    for k in one:
        if k.startswith("approx_"):                       # sums through atomics: order varies
            assert np.max(np.abs(one[k] - grp[k]) / np.abs(one[k]).max()) <= 2e-12 and np.abs(one[k]).max() > 0, k
        elif isinstance(one[k], np.ndarray):
            assert np.array_equal(one[k], grp[k]), k
        else:
            assert one[k] == grp[k], k


@pytest.mark.gpu
def test_restart_per_device_slab_files_gpu(gpu_host, tmp_path, monkeypatch):
    """SURVEY 8f rank 3 on a device group: every device writes its own column block (<file>.rank<r>), and both a
    group and a single device resume from those files bit for bit (omega_hat, SFD state, t, dt, forcing stream)."""
    monkeypatch.setenv("NS_SPECTRAL_GROUP_SAME_DEVICE", "1")
    lib, n = gpu_host, 2048
    path = str(tmp_path / "grp.bin").encode()

    def make(gpus):
        s, fs = _forced_solver(lib, "white", nx=n, ny=n, sfd=True)
        if gpus > 1:
            assert lib.ns_spectral_backend_init(s.sp, gpus) == 0, lib.ns_spectral_last_error()
        lib.ns_spectral_set_sync_policy(s.sp, 1)
        return s, fs

    a, fa = make(2)
    assert lib.ns_spectral_run_steps(a.sp, 3, C.byref(fa)) == 0
    assert lib.ns_spectral_save_state(a.sp, C.byref(fa), path) == 0, lib.ns_spectral_last_error()
    sizes = [os.path.getsize(path.decode() + ".rank%d" % r) for r in range(2)]
    assert all(sz > 16 * n * 300 for sz in sizes)                   # two column blocks, omega_hat + sfd_hat each
    assert os.path.getsize(path.decode()) < 4096                     # the main file holds the header only
    assert lib.ns_spectral_run_steps(a.sp, 3, C.byref(fa)) == 0
    lib.ns_spectral_sync_to_host(a.sp, 0xFFFF)
    want = (np.array(a.arr("omega_hat")), np.array(a.arr("sfd_hat")), a.s.t, a.s.dt, fa.rng_state)
    lib.ns_forcing_state_free(C.byref(fa))
    a.close()
    for gpus in (2, 1):
        b, fb = make(gpus)
        assert lib.ns_spectral_load_state(b.sp, C.byref(fb), path) == 0, lib.ns_spectral_last_error()
        assert lib.ns_spectral_run_steps(b.sp, 3, C.byref(fb)) == 0
        lib.ns_spectral_sync_to_host(b.sp, 0xFFFF)
        assert np.array_equal(b.arr("omega_hat"), want[0]) and np.array_equal(b.arr("sfd_hat"), want[1])
        assert (b.s.t, b.s.dt, fb.rng_state) == want[2:]
        lib.ns_forcing_state_free(C.byref(fb))
        b.close()


@pytest.mark.gpu
def test_cli_runs_with_config_files(cuda_lib, tmp_path):
    """README.md:67-77 usage: repeated --config, later files override."""
    ini = tmp_path / "base.ini"
    ini.write_text("[grid]\nnx = 64\nny = 64\n[fluid]\nRe = 100.0\n[time]\nt_end = 0.3\n[output]\ninterval = 0\n")
    js = tmp_path / "over.json"
    js.write_text('{"grid": {"nx": 128}, "forcing": {"type": "white", "energy_target": 0.5}}')
    exe = os.path.join(ROOT, "navier-stokes-solvers_b200", "bin", "ns_spectral_solver")
    out = subprocess.run([exe, "--config", str(ini), "--config", str(js)], capture_output=True, text=True,
                         cwd=tmp_path, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "Resolution: 128x64 modes" in out.stdout
    assert "=== SIMULATION COMPLETED ===" in out.stdout
    assert "Step 10:" in out.stdout


@pytest.mark.gpu
def test_cli_honours_intervals_that_are_not_multiples_of_ten(cuda_lib, tmp_path):
    """output.interval = 7 and restart.interval = 4: the loop advances in batches but must stop at every
    multiple of both (the reference tests `step % interval == 0` after each step, main_spectral.c:51-64)."""
    ini = tmp_path / "c.ini"
    ini.write_text("[grid]\nnx = 64\nny = 64\n[fluid]\nRe = 100.0\n[time]\nt_end = 100.0\nmax_steps = 23\n"
                   "[output]\ninterval = 7\nprefix = snap\n[restart]\nsave = r.bin\ninterval = 4\n")
    exe = os.path.join(ROOT, "navier-stokes-solvers_b200", "bin", "ns_spectral_solver")
    out = subprocess.run([exe, "--config", str(ini)], capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert out.returncode == 0, out.stderr
    snaps = sorted(p.name for p in tmp_path.glob("snap_step_*.dat"))
    assert snaps == ["snap_step_0007.dat", "snap_step_0014.dat", "snap_step_0021.dat"]
    assert "Step 10:" in out.stdout and "Step 20:" in out.stdout
    assert (tmp_path / "r.bin").exists()


@pytest.mark.gpu
def test_cli_restart_resumes_bit_exact(cuda_lib, tmp_path):
    """restart.save / restart.load: 20 forced steps in one run == 10 + save, load + 10 (same final
    energy and enstrophy lines, printed with 7 significant digits, and the same final time)."""
    exe = os.path.join(ROOT, "navier-stokes-solvers_b200", "bin", "ns_spectral_solver")
    base = ("[grid]\nnx = 64\nny = 64\n[fluid]\nRe = 500.0\n[time]\nt_end = 100.0\nmax_steps = %d\n"
            "[init]\ntype = random_phase\n[forcing]\ntype = colored\nenergy_target = 0.5\nseed = 9\n"
            "[output]\ninterval = 0\n[restart]\n%s\n")
    def run(name, steps, restart):
        f = tmp_path / name
        f.write_text(base % (steps, restart))
        out = subprocess.run([exe, "--config", str(f)], capture_output=True, text=True, cwd=tmp_path, timeout=300)
        assert out.returncode == 0, out.stderr
        return [l for l in out.stdout.split("\n") if l.startswith(("Final time", "Final energy", "Final enstrophy"))]
    whole = run("whole.ini", 20, "save = whole.bin")
    run("first.ini", 10, "save = half.bin")
    second = run("second.ini", 10, "load = half.bin\nsave = second.bin")
    assert len(whole) == 3 and whole == second
    assert (tmp_path / "whole.bin").read_bytes() == (tmp_path / "second.bin").read_bytes()
