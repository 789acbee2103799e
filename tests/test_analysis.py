"""turbulence statistics (include/ns_statistics.h) and VTK export (include/ns_vtk_writer.h)
against the UNMODIFIED reference sources compiled into oracle/_ref (src/analysis/statistics.c,
src/output/vtk_writer.c). The same bodies run on the CPU emulator build (-m "not gpu") and on the product
library (-m gpu)."""
import ctypes as C
import os

import numpy as np
import pytest

import hostlib
import ns_structs as st
import orc

c_double_p = C.POINTER(C.c_double)


class Moments(C.Structure):
    _fields_ = [("mean", C.c_double), ("rms", C.c_double), ("skewness", C.c_double), ("kurtosis", C.c_double)]


class TurbStats(C.Structure):
    _fields_ = [("u", Moments), ("v", Moments), ("integral_length_scale_x", C.c_double),
                ("integral_length_scale_y", C.c_double), ("taylor_microscale", C.c_double),
                ("kolmogorov_scale", C.c_double)]


STAT_ARGS = [c_double_p, c_double_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
             C.POINTER(TurbStats), c_double_p, c_double_p]


def _ref():
    if not orc.have("ref"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    lib = C.CDLL(orc.lib_path("ref"))
    lib.ns_stats_compute_velocity_statistics.argtypes = STAT_ARGS
    return lib


def _flat(s):
    return np.array([s.u.mean, s.u.rms, s.u.skewness, s.u.kurtosis, s.v.mean, s.v.rms, s.v.skewness, s.v.kurtosis,
                     s.integral_length_scale_x, s.integral_length_scale_y, s.taylor_microscale, s.kolmogorov_scale])


def _stats(lib, u, v, dx, dy, nu, sx, sy, with_corr=True):
    ny, nx = u.shape
    out = TurbStats()
    cx, cy = np.full(max(sx, 1), -7.0), np.full(max(sy, 1), -7.0)
    lib.ns_stats_compute_velocity_statistics.argtypes = STAT_ARGS
    rc = lib.ns_stats_compute_velocity_statistics(
        u.ctypes.data_as(c_double_p), v.ctypes.data_as(c_double_p), nx, ny, dx, dy, nu, sx, sy, C.byref(out),
        cx.ctypes.data_as(c_double_p) if with_corr else None, cy.ctypes.data_as(c_double_p) if with_corr else None)
    assert rc == 0
    return _flat(out), cx[:sx], cy[:sy]


def _smooth_field(ny, nx, seed):
    """a periodic field with a few scales so that the correlations cross zero"""
    rng = np.random.default_rng(seed)
    y, x = np.meshgrid(np.arange(ny) * 2 * np.pi / ny, np.arange(nx) * 2 * np.pi / nx, indexing="ij")
    f = np.zeros((ny, nx))
    for _ in range(12):
        kx, ky = rng.integers(1, 7, 2)
        f += rng.standard_normal() * np.cos(kx * x + ky * y + rng.uniform(0, 2 * np.pi))
    return np.ascontiguousarray(f + 0.3 + 0.05 * rng.standard_normal((ny, nx)))


def _check_statistics(lib, shapes):
    ref = _ref()
    for (ny, nx, sx, sy) in shapes:
        u, v = _smooth_field(ny, nx, 1), _smooth_field(ny, nx, 2)
        dx, dy, nu = 2 * np.pi / nx, 2 * np.pi / ny, 1e-3
        want, wcx, wcy = _stats(ref, u, v, dx, dy, nu, sx, sy)
        got, gcx, gcy = _stats(lib, u, v, dx, dy, nu, sx, sy)
        assert np.all(np.abs(got - want) <= 1e-12 * np.maximum(np.abs(want), 1e-3)), (got, want)
        assert np.max(np.abs(gcx - wcx)) <= 1e-12 and np.max(np.abs(gcy - wcy)) <= 1e-12
        assert want[8] > 0 and want[9] > 0 and want[10] > 0 and want[11] > 0
        # optional correlation buffers absent: same scalars
        got2, _, _ = _stats(lib, u, v, dx, dy, nu, sx, sy, with_corr=False)
        assert np.all(np.abs(got2 - want) <= 1e-12 * np.maximum(np.abs(want), 1e-3))
    # separations beyond the grid are clipped and the tail zero-filled (statistics.c:265-266, 306-316)
    u, v = _smooth_field(8, 16, 3), _smooth_field(8, 16, 4)
    want, wcx, wcy = _stats(ref, u, v, 0.1, 0.2, 1e-2, 20, 12)
    got, gcx, gcy = _stats(lib, u, v, 0.1, 0.2, 1e-2, 20, 12)
    assert np.all(np.abs(got - want) <= 1e-12 * np.maximum(np.abs(want), 1e-3))
    assert np.max(np.abs(gcx - wcx)) <= 1e-12 and np.max(np.abs(gcy - wcy)) <= 1e-12 and gcx[16:].max() == 0.0
    # a constant field: zero variance -> delta correlations, zero scales (statistics.c:19-24, 318-324)
    z = np.ones((8, 8))
    want, wcx, _ = _stats(ref, z, z, 0.1, 0.1, 1e-2, 4, 4)
    got, gcx, _ = _stats(lib, z, z, 0.1, 0.1, 1e-2, 4, 4)
    assert np.array_equal(got, want) and np.array_equal(gcx, wcx)
    # invalid input
    bad = TurbStats()
    assert lib.ns_stats_compute_velocity_statistics(None, None, 8, 8, 0.1, 0.1, 1e-2, 0, 0, C.byref(bad), None, None) != 0


def _solver_statistics_and_vtk(lib, tmp_path, n):
    """ns_spectral_turbulence_statistics on the device-resident field (LAZY) == on the host arrays (EAGER) == the
    reference function on the downloaded u, v; VTK files byte-identical to the reference writer's."""
    ref = _ref()
    s = hostlib.Solver(lib, n, n, Re=500.0)
    lib.ns_spectral_initialize_random_phase(s.sp, 5, 6.0, 3.0)
    lib.ns_spectral_turbulence_statistics.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(TurbStats), c_double_p,
                                                      c_double_p]
    sx = sy = 24
    u, v = np.array(s.arr("u")), np.array(s.arr("v"))
    want, wcx, wcy = _stats(ref, u, v, 2 * np.pi / n, 2 * np.pi / n, 1.0 / 500.0, sx, sy)
    for policy in (0, 1):
        lib.ns_spectral_set_sync_policy(s.sp, policy)
        out = TurbStats()
        cx, cy = np.zeros(sx), np.zeros(sy)
        assert lib.ns_spectral_turbulence_statistics(s.sp, sx, sy, C.byref(out), cx.ctypes.data_as(c_double_p),
                                                     cy.ctypes.data_as(c_double_p)) == 0
        got = _flat(out)
        assert np.all(np.abs(got - want) <= 1e-11 * np.maximum(np.abs(want), 1e-3)), (policy, got, want)
        assert np.max(np.abs(cx - wcx)) <= 1e-11 and np.max(np.abs(cy - wcy)) <= 1e-11
    lib.ns_spectral_set_sync_policy(s.sp, 0)          # EAGER: host arrays current again
    # VTK: product writer vs the reference writer on the same struct (layout-compatible prefix)
    for name, fn in [("legacy.vtk", "ns_vtk_write_spectral_legacy"), ("grid.vtr", "ns_vtk_write_spectral_xml")]:
        a, b = tmp_path / ("our_" + name), tmp_path / ("ref_" + name)
        for L, path in ((lib, a), (ref, b)):
            f = getattr(L, fn)
            f.argtypes = [C.c_char_p, C.c_void_p, C.c_double]
            assert f(str(path).encode(), s.sp, 0.125) == 0
        assert a.read_bytes() == b.read_bytes() and a.stat().st_size > 1000
    for L, path in ((lib, tmp_path / "our.pvd"), (ref, tmp_path / "ref.pvd")):
        L.ns_vtk_update_pvd.argtypes = [C.c_char_p, C.c_char_p, C.c_double]
        assert L.ns_vtk_update_pvd(str(path).encode(), b"a.vtr", 0.0) == 0
        assert L.ns_vtk_update_pvd(str(path).encode(), b"b.vtr", 0.5) == 0
    assert (tmp_path / "our.pvd").read_bytes() == (tmp_path / "ref.pvd").read_bytes()
    # strided legacy dump after LAZY steps: the points of the full dump that survive the stride
    lib.ns_spectral_set_sync_policy(s.sp, 1)
    assert lib.ns_spectral_run_steps(s.sp, 2, None) == 0
    lib.ns_spectral_write_vtk_strided.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    full, thin = tmp_path / "full.vtk", tmp_path / "thin.vtk"
    assert lib.ns_spectral_write_vtk_strided(s.sp, str(full).encode(), 1) == 0
    assert lib.ns_spectral_write_vtk_strided(s.sp, str(thin).encode(), 4) == 0
    assert lib.ns_spectral_write_vtk_strided(s.sp, str(thin).encode(), 0) == -1
    fl, tl = full.read_text().split("\n"), thin.read_text().split("\n")
    m = n // 4
    assert tl[4] == f"DIMENSIONS {m} {m} 1" and tl[5] == f"POINTS {m * m} double"
    pick = [j * n + i for j in range(0, n, 4) for i in range(0, n, 4)]
    assert tl[6:6 + m * m] == [fl[6 + k] for k in pick]
    vec0_full, vec0_thin = fl.index("VECTORS velocity double") + 1, tl.index("VECTORS velocity double") + 1
    assert tl[vec0_thin:vec0_thin + m * m] == [fl[vec0_full + k] for k in pick]
    assert f"t={s.s.t:.6f}" in tl[1] and s.s.t > 0
    s.close()


@pytest.fixture(scope="module")
def emu_host(emu_lib):
    return hostlib.load(hostlib.EMU)


@pytest.fixture(scope="module")
def gpu_host(cuda_lib):
    return hostlib.load(hostlib.PRODUCT)


def test_statistics_emu(emu_host):
    _check_statistics(emu_host, [(40, 48, 12, 10), (32, 32, 32, 32)])


def test_solver_statistics_and_vtk_emu(emu_host, tmp_path):
    _solver_statistics_and_vtk(emu_host, tmp_path, 32)


@pytest.mark.gpu
def test_statistics_gpu(gpu_host):
    _check_statistics(gpu_host, [(40, 48, 12, 10), (384, 500, 64, 48), (1024, 1024, 96, 96)])


@pytest.mark.gpu
def test_solver_statistics_and_vtk_gpu(gpu_host, tmp_path):
    _solver_statistics_and_vtk(gpu_host, tmp_path, 256)
