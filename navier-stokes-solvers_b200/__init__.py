"""navier-stokes-solvers_b200 -- B200-native spectral Navier-Stokes time step.

Python is plumbing here: thin ctypes views of the two C-ABI shared libraries

  lib/libnsb.so         kernel ABI  (include/nsb.h; nvcc, sm_100a)
  lib/libns_spectral.so reference-facing entry points (include/ns_spectral_solver.h,
                        include/ns_forcing.h, include/ns_config.h; host C over libnsb)

There is NO CPU fallback: loading fails loudly when the CUDA library is missing and
context creation fails when no GPU is visible. (tests/ may pass an explicit library
path to run the same kernel sources under the CPU emulator; nothing here selects it.)

The directory name contains a hyphen (it mirrors the reference repository's name), so
import it with `importlib` -- see `__graft_entry__.load_package()`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
NSB_LIB = os.path.join(LIB_DIR, "libnsb.so")
HOST_LIB = os.path.join(LIB_DIR, "libns_spectral.so")

# field ids, include/nsb.h
OMEGA_HAT, PSI_HAT, U_HAT, V_HAT, NONLINEAR_HAT, SFD_HAT, RK_ACC, OMEGA_STAGE, SCRATCH_HAT, SCRATCH2_HAT = range(10)
U, V, P, OMEGA, PSI, SCRATCH_PHYS = range(16, 22)

c_double_p = C.POINTER(C.c_double)


class NsbParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("Lx", C.c_double), ("Ly", C.c_double),
                ("Re", C.c_double), ("cfl_factor", C.c_double), ("device", C.c_int),
                ("y1", C.c_int), ("x2", C.c_int), ("col_tile", C.c_int), ("beta_batch", C.c_int)]


class NsbHyperviscosity(C.Structure):
    _fields_ = [("enabled", C.c_int), ("coefficient", C.c_double), ("order", C.c_int),
                ("cutoff_wavenumber", C.c_double), ("inv_bandwidth", C.c_double),
                ("sfd_enabled", C.c_int), ("sfd_strength", C.c_double), ("sfd_timescale", C.c_double)]


class NsbError(RuntimeError):
    pass


def load_nsb(path=None):
    """Open the kernel library. Raises if it has not been built (no fallback)."""
    path = path or NSB_LIB
    if not os.path.exists(path):
        raise NsbError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                       " (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(path)
    lib.nsb_last_error.restype = C.c_char_p
    lib.nsb_ctx_create.argtypes = [C.POINTER(NsbParams), C.POINTER(C.c_void_p)]
    lib.nsb_ctx_destroy.argtypes = [C.c_void_p]
    lib.nsb_ctx_destroy.restype = None
    lib.nsb_stream.argtypes = [C.c_void_p]
    lib.nsb_stream.restype = C.c_void_p
    lib.nsb_launch_count.argtypes = [C.c_void_p]
    lib.nsb_launch_count.restype = C.c_longlong
    sigs = {
        "nsb_sync": [C.c_void_p],
        "nsb_upload": [C.c_void_p, C.c_int, C.c_void_p],
        "nsb_download": [C.c_void_p, C.c_int, C.c_void_p],
        "nsb_zero": [C.c_void_p, C.c_int],
        "nsb_copy": [C.c_void_p, C.c_int, C.c_int],
        "nsb_set_hyperviscosity": [C.c_void_p, C.POINTER(NsbHyperviscosity)],
        "nsb_set_forcing": [C.c_void_p, C.c_int, C.c_int, C.c_void_p],
        "nsb_physical_to_spectral": [C.c_void_p, C.c_int, C.c_int],
        "nsb_spectral_to_physical": [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)],
        "nsb_apply_dealiasing": [C.c_void_p, C.c_int],
        "nsb_velocity_from_vorticity": [C.c_void_p],
        "nsb_compute_derivatives": [C.c_void_p],
        "nsb_nonlinear_terms": [C.c_void_p],
        "nsb_hyperviscosity_apply": [C.c_void_p, C.c_int, C.c_int],
        "nsb_hyperviscosity_post_step": [C.c_void_p, C.c_int, C.c_double],
        "nsb_cfl_timestep_uv": [C.c_void_p, c_double_p],
        "nsb_cfl_timestep": [C.c_void_p, c_double_p],
        "nsb_rk4_step": [C.c_void_p, C.c_double],
        "nsb_diagnostics": [C.c_void_p, c_double_p],
        "nsb_diagnostics_phys": [C.c_void_p, c_double_p],
        "nsb_get_state": [C.c_void_p, c_double_p, c_double_p, C.POINTER(C.c_int)],
        "nsb_set_time": [C.c_void_p, C.c_double],
        "nsb_get_dt_history": [C.c_void_p, C.c_int, c_double_p],
        "nsb_get_plan": [C.c_void_p, C.POINTER(C.c_int)],
        "nsb_profile_begin": [C.c_void_p],
        "nsb_profile_end": [C.c_void_p, c_double_p, C.POINTER(C.c_int)],
        "nsb_timer_start": [C.c_void_p],
        "nsb_timer_stop": [C.c_void_p, C.POINTER(C.c_float)],
        "nsb_mgpu_setup": [C.c_void_p, C.c_int, C.c_int],
        "nsb_mgpu_partition": [C.c_void_p, C.POINTER(C.c_int)],
        "nsb_mgpu_export": [C.c_void_p, C.c_char_p],
        "nsb_mgpu_import": [C.c_void_p, C.c_int, C.c_char_p],
        "nsb_mgpu_link_local": [C.POINTER(C.c_void_p), C.c_int],
        "nsb_mgpu_group_step": [C.POINTER(C.c_void_p), C.c_int, C.c_double],
        "nsb_spectral_stats": [C.c_void_p, c_double_p],
        "nsb_energy_spectrum": [C.c_void_p, C.c_int, c_double_p, C.POINTER(C.c_int)],
        "nsb_mgpu_group_gather": [C.POINTER(C.c_void_p), C.c_int],
        "nsb_state_checksum": [C.c_void_p, C.POINTER(C.c_ulonglong)],
        "nsb_diagnostics_spectral": [C.c_void_p, c_double_p],
        "nsb_mgpu_group_scatter": [C.POINTER(C.c_void_p), C.c_int],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.nsb_device_count.restype = C.c_int
    return lib


class NsbContext:
    """One device-resident solver instance over the kernel ABI (include/nsb.h)."""

    def __init__(self, nx, ny, Lx=2 * np.pi, Ly=2 * np.pi, Re=100.0, cfl=0.3, device=0,
                 y1=0, x2=0, col_tile=0, beta_batch=0, lib=None):
        self.lib = lib if lib is not None else load_nsb()
        self.nx, self.ny, self.nkx = nx, ny, nx // 2 + 1
        self.Lx, self.Ly, self.Re = Lx, Ly, Re
        p = NsbParams(nx, ny, Lx, Ly, Re, cfl, device, y1, x2, col_tile, beta_batch)
        h = C.c_void_p()
        if self.lib.nsb_ctx_create(C.byref(p), C.byref(h)) != 0:
            raise NsbError(self.lib.nsb_last_error().decode())
        self.h = h
        self._force_geom = None

    def _ck(self, rc):
        if rc != 0:
            raise NsbError(self.lib.nsb_last_error().decode())

    def close(self):
        if self.h:
            self.lib.nsb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- data movement ----
    def _shape(self, field):
        return (self.ny, self.nx) if field >= 16 else (self.ny, self.nkx)

    def upload(self, field, arr):
        dt = np.float64 if field >= 16 else np.complex128
        a = np.ascontiguousarray(arr, dtype=dt)
        assert a.shape == self._shape(field), (a.shape, self._shape(field))
        self._ck(self.lib.nsb_upload(self.h, field, a.ctypes.data_as(C.c_void_p)))

    def download(self, field):
        dt = np.float64 if field >= 16 else np.complex128
        a = np.empty(self._shape(field), dtype=dt)
        self._ck(self.lib.nsb_download(self.h, field, a.ctypes.data_as(C.c_void_p)))
        return a

    def sync(self):
        self._ck(self.lib.nsb_sync(self.h))

    # ---- configuration ----
    def k2(self):
        kx = 2.0 * np.pi * np.arange(self.nkx) / self.Lx
        j = np.arange(self.ny)
        ky = np.where(j <= self.ny // 2, 2.0 * np.pi * j / self.Ly, 2.0 * np.pi * (j - self.ny) / self.Ly)
        return kx[None, :] * kx[None, :] + ky[:, None] * ky[:, None]

    def set_hyperviscosity(self, enabled=1, coefficient=0.0, order=8, cutoff_ratio=0.65,
                           sfd_enabled=0, sfd_strength=0.0, sfd_timescale=1.0):
        """Derives kmax / cutoff / inv_bandwidth like hyperviscosity.c:68-78."""
        kmax = float(np.sqrt(self.k2()).max())
        kc = min(max(cutoff_ratio, 0.0), 1.0) * kmax
        bw = max(kmax - kc, 1e-12)
        sfd_on = bool(sfd_enabled) and sfd_strength > 0.0
        h = NsbHyperviscosity(int(bool(enabled) and coefficient > 0.0), coefficient, max(order, 1), kc,
                              1.0 / bw, int(sfd_on), sfd_strength if sfd_on else 0.0, sfd_timescale)
        self._ck(self.lib.nsb_set_hyperviscosity(self.h, C.byref(h)))
        if sfd_on:
            self._ck(self.lib.nsb_copy(self.h, SFD_HAT, OMEGA_HAT))

    def set_forcing_dense(self, f_hat):
        """f_hat: full (ny, nkx) complex array (mostly zeros); packs its bounding box."""
        if f_hat is None:
            self._ck(self.lib.nsb_set_forcing(self.h, 0, 0, None))
            return
        if self._force_geom is None:
            nz = np.argwhere(f_hat != 0)
            if nz.size == 0:
                ib = jb = 0
            else:
                j = nz[:, 0]
                band = np.where(j <= self.ny // 2, j, j - self.ny)
                ib, jb = int(nz[:, 1].max()), int(np.abs(band).max())
            self._force_geom = (ib, jb)
        ib, jb = self._force_geom
        bands = np.arange(-jb, jb + 1)
        box = np.ascontiguousarray(f_hat[bands % self.ny][:, : ib + 1])
        self._ck(self.lib.nsb_set_forcing(self.h, ib, jb, box.ctypes.data_as(C.c_void_p)))

    # ---- operations ----
    def physical_to_spectral(self, src, dst):
        self._ck(self.lib.nsb_physical_to_spectral(self.h, src, dst))

    def spectral_to_physical(self, srcs, dsts):
        n = len(srcs)
        a = (C.c_int * n)(*srcs)
        b = (C.c_int * n)(*dsts)
        self._ck(self.lib.nsb_spectral_to_physical(self.h, n, a, b))

    def apply_dealiasing(self, field):
        self._ck(self.lib.nsb_apply_dealiasing(self.h, field))

    def velocity_from_vorticity(self):
        self._ck(self.lib.nsb_velocity_from_vorticity(self.h))

    def compute_derivatives(self):
        self._ck(self.lib.nsb_compute_derivatives(self.h))

    def nonlinear_terms(self):
        self._ck(self.lib.nsb_nonlinear_terms(self.h))

    def cfl_timestep(self):
        d = C.c_double()
        self._ck(self.lib.nsb_cfl_timestep(self.h, C.byref(d)))
        return d.value

    def step(self, dt=0.0):
        self._ck(self.lib.nsb_rk4_step(self.h, float(dt)))

    def diagnostics(self):
        out = (C.c_double * 4)()
        self._ck(self.lib.nsb_diagnostics(self.h, out))
        n = float(self.nx * self.ny)
        return {"energy": 0.5 * (out[0] + out[1]) / n, "enstrophy": 0.5 * out[2] / n,
                "max_speed": float(np.sqrt(out[3]))}

    def diagnostics_spectral(self):
        """energy / enstrophy by Parseval over the retained modes (no transforms)"""
        out = (C.c_double * 3)()
        self._ck(self.lib.nsb_diagnostics_spectral(self.h, out))
        n = float(self.nx * self.ny)
        return {"energy": 0.5 * (out[0] + out[1]) / n, "enstrophy": 0.5 * out[2] / n}

    def spectral_stats(self):
        out = (C.c_double * 6)()
        self._ck(self.lib.nsb_spectral_stats(self.h, out))
        return dict(zip(["total", "high_k", "max_abs", "tail", "tail_modes", "k2_last"], out[:]))

    def energy_spectrum(self, nbins):
        """shell sums of 0.5 (|u_hat|^2 + |v_hat|^2) and mode counts (spectral_analysis.c:103-142)"""
        ek, cnt = (C.c_double * nbins)(), (C.c_int * nbins)()
        self._ck(self.lib.nsb_energy_spectrum(self.h, nbins, ek, cnt))
        return np.array(ek[:]), np.array(cnt[:])

    def state_checksum(self):
        """(xor, wrapping sum) of the bit patterns of omega_hat over the own retained columns"""
        out = (C.c_ulonglong * 2)()
        self._ck(self.lib.nsb_state_checksum(self.h, out))
        return int(out[0]), int(out[1])

    def state(self):
        dt, t, st = C.c_double(), C.c_double(), C.c_int()
        self._ck(self.lib.nsb_get_state(self.h, C.byref(dt), C.byref(t), C.byref(st)))
        return dt.value, t.value, st.value

    def dt_history(self, n):
        out = (C.c_double * n)()
        self._ck(self.lib.nsb_get_dt_history(self.h, n, out))
        return np.array(out[:])

    def plan(self):
        out = (C.c_int * 8)()
        self._ck(self.lib.nsb_get_plan(self.h, out))
        return dict(zip(["Y1", "Y2", "X1", "X2", "col_tile", "beta_batch", "ncol", "launches_per_step"], out[:]))

    def timer_start(self):
        self._ck(self.lib.nsb_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self.lib.nsb_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    # ---- slab decomposition (include/nsb.h: nsb_mgpu_*) ----
    def mgpu_setup(self, rank, world):
        self._ck(self.lib.nsb_mgpu_setup(self.h, rank, world))

    def partition(self):
        out = (C.c_int * 6)()
        self._ck(self.lib.nsb_mgpu_partition(self.h, out))
        return dict(zip(["rank", "world", "col0", "col1", "row0", "rows"], out[:]))

    def mgpu_export(self):
        buf = C.create_string_buffer(192)
        self._ck(self.lib.nsb_mgpu_export(self.h, buf))
        return buf.raw

    def mgpu_import(self, peer, handles):
        self._ck(self.lib.nsb_mgpu_import(self.h, peer, handles))

    def launch_count(self):
        return int(self.lib.nsb_launch_count(self.h))

    def stream(self):
        return self.lib.nsb_stream(self.h)


class NsbGroup:
    """One process driving `world` devices: slab-decomposed contexts linked through peer memory
    (nsb_mgpu_link_local / nsb_mgpu_group_step)."""

    def __init__(self, nx, ny, world, devices=None, lib=None, **kw):
        self.lib = lib if lib is not None else load_nsb()
        devices = devices if devices is not None else list(range(world))
        self.ctx = [NsbContext(nx, ny, device=devices[r], lib=self.lib, **kw) for r in range(world)]
        for r, c in enumerate(self.ctx):
            c.mgpu_setup(r, world)
        self._arr = (C.c_void_p * world)(*[c.h for c in self.ctx])
        if self.lib.nsb_mgpu_link_local(self._arr, world) != 0:
            raise NsbError(self.lib.nsb_last_error().decode())
        self.world = world

    def each(self, fn):
        return [fn(c) for c in self.ctx]

    def upload(self, field, arr):
        self.each(lambda c: c.upload(field, arr))

    def step(self, dt=0.0):
        if self.lib.nsb_mgpu_group_step(self._arr, self.world, float(dt)) != 0:
            raise NsbError(self.lib.nsb_last_error().decode())

    def sync(self):
        self.each(lambda c: c.sync())

    def download_omega_hat(self):
        """Merge the own-column blocks of all devices (masked / unowned modes from rank 0)."""
        out = None
        for c in self.ctx:
            a = c.download(OMEGA_HAT)
            p = c.partition()
            if out is None:
                out = a
            else:
                out[:, p["col0"]:p["col1"]] = a[:, p["col0"]:p["col1"]]
        return out

    def close(self):
        self.each(lambda c: c.close())
