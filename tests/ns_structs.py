"""ctypes descriptions of the PUBLIC C headers (include/ns_spectral_solver.h, include/ns_forcing.h):
struct layouts and numpy views. Shared by the product-side harness (tests/hostlib.py, bench.py) and
by the oracle harness (tests/orc.py); contains no oracle code and loads no library."""
import ctypes as C

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class Hyperviscosity(C.Structure):
    _fields_ = [
        ("enabled", C.c_int),
        ("coefficient", C.c_double),
        ("order", C.c_int),
        ("cutoff_ratio", C.c_double),
        ("sfd_strength", C.c_double),
        ("sfd_timescale", C.c_double),
        ("sfd_enabled", C.c_int),
        ("kmax", C.c_double),
        ("cutoff_wavenumber", C.c_double),
        ("inv_bandwidth", C.c_double),
    ]


REAL_FIELDS = ["u", "v", "p", "omega", "psi"]
CPLX_FIELDS = ["u_hat", "v_hat", "omega_hat", "psi_hat", "nonlinear_hat", "rk1_hat", "rk2_hat",
               "rk3_hat", "rk4_hat", "omega_tmp_hat"]

_SPECTRAL_FIELDS = (
    [("nx", C.c_int), ("ny", C.c_int), ("nkx", C.c_int), ("nky", C.c_int),
     ("Lx", C.c_double), ("Ly", C.c_double), ("Re", C.c_double), ("dt", C.c_double),
     ("t", C.c_double), ("cfl_factor", C.c_double)]
    + [(n, c_double_p) for n in REAL_FIELDS + ["x", "y"]]
    + [(n, c_double_p) for n in CPLX_FIELDS]
    + [("kx", c_double_p), ("ky", c_double_p), ("k2", c_double_p),
       ("forward_plan", C.c_void_p), ("backward_plan", C.c_void_p),
       ("dealias_mask", c_int_p),
       ("hyperviscosity", Hyperviscosity),
       ("hyperviscosity_weights", c_double_p),
       ("sfd_hat", c_double_p)]
)


class SpectralStruct(C.Structure):
    _fields_ = _SPECTRAL_FIELDS


class ForcingConfig(C.Structure):
    _fields_ = [("kmin", C.c_double), ("kmax", C.c_double), ("energy_target", C.c_double),
                ("correlation_time", C.c_double), ("seed", C.c_ulong), ("type", C.c_int)]


class ForcingState(C.Structure):
    _fields_ = [("config", ForcingConfig), ("nkx", C.c_int), ("nky", C.c_int),
                ("spectral_size", C.c_int), ("active_modes", C.c_int),
                ("weights", c_double_p), ("colored_state", c_double_p),
                ("have_spare", C.c_uint), ("spare_gaussian", C.c_double),
                ("rng_state", C.c_ulonglong)]


class View:
    """numpy views onto the arrays of a (reference-layout) spectral struct."""

    def __init__(self, sp):
        self.sp = sp
        s = sp.contents
        self.nx, self.ny, self.nkx, self.nky = s.nx, s.ny, s.nkx, s.nky

    def real(self, name):
        s = self.sp.contents
        return np.ctypeslib.as_array(getattr(s, name), shape=(self.ny, self.nx))

    def cplx(self, name):
        s = self.sp.contents
        a = np.ctypeslib.as_array(getattr(s, name), shape=(self.nky, self.nkx * 2))
        return a.view(np.complex128)

    def vec(self, name, n):
        return np.ctypeslib.as_array(getattr(self.sp.contents, name), shape=(n,))

    @property
    def k2(self):
        return np.ctypeslib.as_array(self.sp.contents.k2, shape=(self.nky, self.nkx))

    @property
    def mask(self):
        return np.ctypeslib.as_array(self.sp.contents.dealias_mask, shape=(self.nky, self.nkx))

    @property
    def weights(self):
        return np.ctypeslib.as_array(self.sp.contents.hyperviscosity_weights,
                                     shape=(self.nky, self.nkx))
