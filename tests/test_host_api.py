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


def test_reference_unit_tests_emu(emu_host):
    _reference_unit_tests(emu_host)


def test_conservation_and_eager_parity_emu(emu_host, oracles):
    _conservation_tests(emu_host, oracles, steps=6)


@pytest.mark.parametrize("kind", ["white", "colored"])
def test_lazy_run_steps_emu(emu_host, oracles, kind):
    _lazy_run_steps(emu_host, oracles, kind)


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
