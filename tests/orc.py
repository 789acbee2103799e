"""ctypes access to the two CPU oracles (TEST INFRASTRUCTURE ONLY).

  kind="ref"  -> oracle/_ref/libns_ref.so : the UNMODIFIED reference sources
                 (/root/reference/src/spectral/*.c, src/common/forcing.c) built
                 against oracle/shim/fftw3.h, plus oracle/ref_driver.c
  kind="port" -> oracle/libns_oracle.so   : the C restatement oracle/ns_oracle.c

Both expose the same struct layout (src/spectral/ns_spectral_solver.h:22-68) so a
single ctypes description serves them -- and the product, which appends one
opaque pointer.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

from ns_structs import (CPLX_FIELDS, REAL_FIELDS, ForcingConfig, ForcingState, Hyperviscosity,  # noqa: F401
                        SpectralStruct, View, _SPECTRAL_FIELDS, c_double_p, c_int_p)


def build_oracles():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True,
                   stdout=subprocess.DEVNULL)


def lib_path(kind, flavour="parity"):
    if kind == "ref":
        v4 = os.path.join(ORACLE_DIR, "_ref", "libns_ref_v4.so")
        if flavour == "timing" and os.path.exists(v4) and _cpu_has("avx512f"):
            return v4
        return os.path.join(ORACLE_DIR, "_ref", "libns_ref.so")
    return os.path.join(ORACLE_DIR, "libns_oracle.so")


def _cpu_has(flag):
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return flag in line.split()
    except OSError:
        pass
    return False


def build_flavour(kind, flavour="timing"):
    """compiler flags of the library lib_path() selects (for the bench line's provenance text)"""
    if kind == "ref" and lib_path(kind, flavour).endswith("_v4.so"):
        return "gcc -O3 -fopenmp -march=x86-64-v4 -ffp-contract=off"
    return "gcc -O3 -fopenmp -mavx2 -mfma -ffp-contract=off"


def set_threads(n):
    """OMP threads of the oracle libraries. OMP_NUM_THREADS is read once when libgomp initialises, which
    torch / numpy may already have triggered (and torchrun exports OMP_NUM_THREADS=1), so set it through
    the runtime as well. Returns the number in effect."""
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        gomp = C.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(int(n))
        return int(gomp.omp_get_max_threads())
    except OSError:
        return int(n)


def have(kind):
    return os.path.exists(lib_path(kind))


class Oracle:
    """One solver instance on one of the CPU oracles."""

    def __init__(self, kind, nx, ny, Lx=2 * np.pi, Ly=2 * np.pi, Re=100.0, threads=None, flavour="parity"):
        self.kind = kind
        self.lib = C.CDLL(lib_path(kind, flavour))
        self.pfx = "ns_spectral_" if kind == "ref" else "orc_spectral_"
        self.fpfx = "ns_forcing_" if kind == "ref" else "orc_forcing_"
        L = self.lib
        SP = C.POINTER(SpectralStruct)
        f = self._fn
        f("allocate", SP, [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double])
        for name in ["free", "initialize_grids", "initialize_fft_plans", "initialize_taylor_green",
                     "compute_derivatives", "compute_velocity_from_vorticity",
                     "compute_nonlinear_terms", "advance_rk4", "hyperviscosity_update_weights"]:
            f(name, None, [SP])
        for name in ["compute_cfl_timestep", "compute_energy", "compute_enstrophy"]:
            f(name, C.c_double, [SP])
        f("physical_to_spectral", None, [SP, c_double_p, c_double_p])
        f("spectral_to_physical", None, [SP, c_double_p, c_double_p])
        f("apply_dealiasing", None, [SP, c_double_p])
        f("hyperviscosity_default", None, [C.POINTER(Hyperviscosity)])
        f("hyperviscosity_set", None, [SP, C.POINTER(Hyperviscosity)])
        f("hyperviscosity_apply", None, [SP, c_double_p, c_double_p])
        f("hyperviscosity_post_step", None, [SP, c_double_p, C.c_double])
        self._forced = getattr(L, "nsref_forced_rk4" if kind == "ref" else "orc_forced_rk4")
        self._forced.restype = None
        self._forced.argtypes = [SP, c_double_p]
        self._rp = getattr(L, "nsref_random_phase_init" if kind == "ref" else "orc_random_phase_init")
        self._rp.restype = None
        self._rp.argtypes = [SP, C.c_ulonglong, C.c_double, C.c_double]
        for n, res, args in [
            ("config_default", None, [C.POINTER(ForcingConfig)]),
            ("state_init", C.c_int, [C.POINTER(ForcingState), C.POINTER(ForcingConfig), c_double_p,
                                     C.c_int, C.c_int]),
            ("state_free", None, [C.POINTER(ForcingState)]),
            ("generate", None, [C.POINTER(ForcingState), C.c_double, c_double_p, c_double_p]),
        ]:
            fn = getattr(L, self.fpfx + n)
            fn.restype = res
            fn.argtypes = args
            setattr(self, "f_" + n, fn)
        if threads:
            os.environ.setdefault("OMP_NUM_THREADS", str(threads))
        self.sp = self.allocate(nx, ny, Lx, Ly, Re)
        if not self.sp:
            raise MemoryError("oracle allocate failed")
        if kind == "ref":
            z = L.nsref_zero_fill
            z.restype = None
            z.argtypes = [SP]
            z(self.sp)
        self.initialize_grids(self.sp)
        self.initialize_fft_plans(self.sp)
        if not self.sp.contents.forward_plan:
            raise ValueError("oracle FFT supports power-of-two sizes only")
        self.view = View(self.sp)
        self.forcing_state = None
        self._forcing_buf = None

    def _fn(self, name, res, args):
        fn = getattr(self.lib, self.pfx + name)
        fn.restype = res
        fn.argtypes = args
        setattr(self, name, fn)

    # ---- convenience ----
    @property
    def s(self):
        return self.sp.contents

    def arr(self, name):
        if name in REAL_FIELDS:
            return self.view.real(name)
        return self.view.cplx(name)

    def taylor_green(self):
        self.initialize_taylor_green(self.sp)
        self.compute_derivatives(self.sp)   # canonical sequence, SURVEY G7

    def random_phase(self, seed=7, k0=8.0, max_omega=5.0):
        self._rp(self.sp, seed, k0, max_omega)

    def set_omega_hat(self, w):
        self.arr("omega_hat")[...] = w
        self.compute_derivatives(self.sp)

    def set_hyperviscosity(self, enabled=1, coefficient=0.0, order=8, cutoff_ratio=0.65,
                           sfd_enabled=0, sfd_strength=0.0, sfd_timescale=1.0):
        h = Hyperviscosity()
        self.hyperviscosity_default(C.byref(h))
        h.enabled, h.coefficient, h.order, h.cutoff_ratio = enabled, coefficient, order, cutoff_ratio
        h.sfd_enabled, h.sfd_strength, h.sfd_timescale = sfd_enabled, sfd_strength, sfd_timescale
        self.hyperviscosity_set(self.sp, C.byref(h))

    def init_forcing(self, kind="white", kmin=0.5, kmax=4.0, energy_target=1.25, tau=0.4, seed=42):
        cfg = ForcingConfig()
        self.f_config_default(C.byref(cfg))
        cfg.kmin, cfg.kmax, cfg.energy_target = kmin, kmax, energy_target
        cfg.correlation_time, cfg.seed = tau, seed
        cfg.type = 1 if kind == "colored" else 0
        self.forcing_state = ForcingState()
        rc = self.f_state_init(C.byref(self.forcing_state), C.byref(cfg), self.s.k2,
                               self.s.nkx, self.s.nky)
        if rc != 0:
            raise ValueError("forcing band has no active mode")
        self._forcing_buf = np.zeros((self.s.nky, self.s.nkx), dtype=np.complex128)

    def gen_forcing(self, dt):
        buf = self._forcing_buf
        self.f_generate(C.byref(self.forcing_state), dt, self.s.k2,
                        buf.ctypes.data_as(c_double_p))
        return buf

    def step(self, dt=None, forcing=None):
        """dt = cfl unless given; one (forced) RK4 step. Returns the dt used."""
        if dt is None:
            dt = self.compute_cfl_timestep(self.sp)
        self.s.dt = dt
        if forcing is None and self.forcing_state is not None:
            forcing = self.gen_forcing(dt)
        if forcing is None:
            self.advance_rk4(self.sp)
        else:
            self._forced(self.sp, forcing.ctypes.data_as(c_double_p))
        return dt

    def energy(self):
        return self.compute_energy(self.sp)

    def enstrophy(self):
        return self.compute_enstrophy(self.sp)

    def close(self):
        if self.forcing_state is not None:
            self.f_state_free(C.byref(self.forcing_state))
            self.forcing_state = None
        if self.sp:
            self.free(self.sp)
            self.sp = None
