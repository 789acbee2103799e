"""ctypes view of the reference-facing host library (include/ns_spectral_solver.h,
ns_forcing.h, ns_config.h): product build (lib/libns_spectral.so, needs a GPU) or the
emulator-linked build used by CPU tests (tests/emu/build/libns_spectral_emu.so)."""
import ctypes as C
import os

import numpy as np

import ns_structs as orc  # struct layouts / views of the public headers only (no oracle code)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT = os.path.join(ROOT, "navier-stokes-solvers_b200", "lib", "libns_spectral.so")
EMU = os.path.join(ROOT, "tests", "emu", "build", "libns_spectral_emu.so")

c_double_p = C.POINTER(C.c_double)


class ProductStruct(C.Structure):
    """NSSpectralData of include/ns_spectral_solver.h = reference layout + `backend`."""
    _fields_ = orc._SPECTRAL_FIELDS + [("backend", C.c_void_p)]


class ConfigEntry(C.Structure):
    _fields_ = [("key", C.c_char_p), ("value", C.c_char_p)]


class Config(C.Structure):
    _fields_ = [("entries", C.POINTER(ConfigEntry)), ("count", C.c_size_t), ("capacity", C.c_size_t)]


class SchemaEntry(C.Structure):
    _fields_ = [("key", C.c_char_p), ("type", C.c_int), ("required", C.c_int), ("min_value", C.c_double),
                ("max_value", C.c_double), ("has_min", C.c_int), ("has_max", C.c_int),
                ("allowed_values", C.POINTER(C.c_char_p)), ("allowed_count", C.c_size_t)]


SPECTRAL_SYMBOLS = [
    "ns_spectral_allocate", "ns_spectral_free", "ns_spectral_initialize_grids",
    "ns_spectral_initialize_fft_plans", "ns_spectral_initialize_taylor_green",
    "ns_spectral_physical_to_spectral", "ns_spectral_spectral_to_physical", "ns_spectral_apply_dealiasing",
    "ns_spectral_compute_derivatives", "ns_spectral_compute_velocity_from_vorticity",
    "ns_spectral_compute_nonlinear_terms", "ns_spectral_advance_rk4", "ns_spectral_compute_cfl_timestep",
    "ns_spectral_compute_energy", "ns_spectral_compute_enstrophy", "ns_spectral_analyze_convergence",
    "ns_spectral_analyze_spectrum", "ns_spectral_check_resolution", "ns_spectral_hyperviscosity_default",
    "ns_spectral_hyperviscosity_set", "ns_spectral_hyperviscosity_update_weights",
    "ns_spectral_hyperviscosity_apply", "ns_spectral_hyperviscosity_post_step", "ns_spectral_output_solution",
    # additive
    "ns_spectral_backend_select_device", "ns_spectral_set_sync_policy", "ns_spectral_get_sync_policy",
    "ns_spectral_sync_to_host", "ns_spectral_sync_to_device", "ns_spectral_advance_rk4_forced",
    "ns_spectral_run_steps", "ns_spectral_initialize_random_phase", "ns_spectral_last_error",
    "ns_spectral_backend_context", "ns_spectral_backend_init",
    "ns_spectral_output_solution_strided", "ns_spectral_energy_spectrum", "ns_spectral_save_state",
    "ns_spectral_load_state",
]
ANALYSIS_SYMBOLS = ["ns_stats_compute_velocity_statistics", "ns_spectral_turbulence_statistics",
                    "ns_vtk_write_spectral_legacy", "ns_vtk_write_spectral_xml", "ns_vtk_update_pvd",
                    "ns_spectral_write_vtk_strided"]
FORCING_SYMBOLS = ["ns_forcing_config_default", "ns_forcing_state_init", "ns_forcing_state_free",
                   "ns_forcing_generate", "ns_forcing_box_extent", "ns_forcing_active_indices",
                   "ns_forcing_generate_box"]
CONFIG_SYMBOLS = ["ns_config_init", "ns_config_free", "ns_config_set", "ns_config_get_string",
                  "ns_config_get_int", "ns_config_get_double", "ns_config_get_bool", "ns_config_has_key",
                  "ns_config_load_files", "ns_config_validate"]
NSB_SYMBOLS = [
    "nsb_last_error", "nsb_device_count", "nsb_ctx_create", "nsb_ctx_destroy", "nsb_sync", "nsb_upload",
    "nsb_download", "nsb_zero", "nsb_copy", "nsb_set_hyperviscosity", "nsb_set_forcing",
    "nsb_physical_to_spectral", "nsb_spectral_to_physical", "nsb_apply_dealiasing",
    "nsb_velocity_from_vorticity", "nsb_compute_derivatives", "nsb_nonlinear_terms",
    "nsb_hyperviscosity_apply", "nsb_hyperviscosity_post_step", "nsb_cfl_timestep_uv", "nsb_cfl_timestep",
    "nsb_rk4_step", "nsb_diagnostics", "nsb_diagnostics_phys", "nsb_get_state", "nsb_set_time",
    "nsb_get_dt_history", "nsb_get_plan", "nsb_stream", "nsb_timer_start", "nsb_timer_stop", "nsb_profile_begin", "nsb_profile_end",
    "nsb_launch_count",
    "nsb_mgpu_setup", "nsb_mgpu_partition", "nsb_mgpu_export", "nsb_mgpu_import", "nsb_mgpu_link_local",
    "nsb_mgpu_group_step", "nsb_mgpu_group_gather", "nsb_mgpu_group_scatter", "nsb_spectral_stats",
    "nsb_energy_spectrum", "nsb_state_checksum", "nsb_phys_absmax", "nsb_phys_scale", "nsb_set_physics", "nsb_velocity_statistics",
    "nsb_velocity_statistics_host", "nsb_check_finite", "nsb_download_columns", "nsb_rk4_step_begin", "nsb_rk4_step_finish", "nsb_diagnostics_spectral",
]


def load(path):
    lib = C.CDLL(path)
    SP = C.POINTER(ProductStruct)
    lib.ns_spectral_allocate.restype = SP
    lib.ns_spectral_allocate.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
    for n in ["free", "initialize_grids", "initialize_fft_plans", "initialize_taylor_green",
              "compute_derivatives", "compute_velocity_from_vorticity", "compute_nonlinear_terms",
              "advance_rk4", "hyperviscosity_update_weights", "analyze_spectrum"]:
        f = getattr(lib, "ns_spectral_" + n)
        f.restype = None
        f.argtypes = [SP]
    for n in ["compute_cfl_timestep", "compute_energy", "compute_enstrophy"]:
        f = getattr(lib, "ns_spectral_" + n)
        f.restype = C.c_double
        f.argtypes = [SP]
    lib.ns_spectral_check_resolution.restype = C.c_int
    lib.ns_spectral_check_resolution.argtypes = [SP]
    lib.ns_spectral_analyze_convergence.restype = None
    lib.ns_spectral_analyze_convergence.argtypes = [SP, C.c_int]
    lib.ns_spectral_physical_to_spectral.restype = None
    lib.ns_spectral_physical_to_spectral.argtypes = [SP, c_double_p, c_double_p]
    lib.ns_spectral_spectral_to_physical.restype = None
    lib.ns_spectral_spectral_to_physical.argtypes = [SP, c_double_p, c_double_p]
    lib.ns_spectral_apply_dealiasing.restype = None
    lib.ns_spectral_apply_dealiasing.argtypes = [SP, c_double_p]
    lib.ns_spectral_hyperviscosity_default.restype = None
    lib.ns_spectral_hyperviscosity_default.argtypes = [C.POINTER(orc.Hyperviscosity)]
    lib.ns_spectral_hyperviscosity_set.restype = None
    lib.ns_spectral_hyperviscosity_set.argtypes = [SP, C.POINTER(orc.Hyperviscosity)]
    lib.ns_spectral_hyperviscosity_apply.restype = None
    lib.ns_spectral_hyperviscosity_apply.argtypes = [SP, c_double_p, c_double_p]
    lib.ns_spectral_hyperviscosity_post_step.restype = None
    lib.ns_spectral_hyperviscosity_post_step.argtypes = [SP, c_double_p, C.c_double]
    lib.ns_spectral_output_solution.restype = None
    lib.ns_spectral_output_solution.argtypes = [SP, C.c_char_p]
    lib.ns_spectral_set_sync_policy.restype = None
    lib.ns_spectral_set_sync_policy.argtypes = [SP, C.c_int]
    lib.ns_spectral_sync_to_host.argtypes = [SP, C.c_uint]
    lib.ns_spectral_sync_to_device.argtypes = [SP, C.c_uint]
    lib.ns_spectral_advance_rk4_forced.argtypes = [SP, c_double_p]
    lib.ns_spectral_run_steps.argtypes = [SP, C.c_int, C.POINTER(orc.ForcingState)]
    lib.ns_spectral_initialize_random_phase.argtypes = [SP, C.c_ulonglong, C.c_double, C.c_double]
    lib.ns_spectral_last_error.restype = C.c_char_p
    lib.ns_spectral_backend_init.argtypes = [SP, C.c_int]
    lib.ns_spectral_output_solution_strided.argtypes = [SP, C.c_char_p, C.c_int]
    lib.ns_spectral_energy_spectrum.argtypes = [SP, C.c_int, c_double_p, C.POINTER(C.c_int)]
    lib.ns_spectral_save_state.argtypes = [SP, C.POINTER(orc.ForcingState), C.c_char_p]
    lib.ns_spectral_load_state.argtypes = [SP, C.POINTER(orc.ForcingState), C.c_char_p]
    lib.ns_forcing_config_default.restype = None
    lib.ns_forcing_config_default.argtypes = [C.POINTER(orc.ForcingConfig)]
    lib.ns_forcing_state_init.argtypes = [C.POINTER(orc.ForcingState), C.POINTER(orc.ForcingConfig), c_double_p,
                                          C.c_int, C.c_int]
    lib.ns_forcing_state_free.restype = None
    lib.ns_forcing_state_free.argtypes = [C.POINTER(orc.ForcingState)]
    lib.ns_forcing_generate.restype = None
    lib.ns_forcing_generate.argtypes = [C.POINTER(orc.ForcingState), C.c_double, c_double_p, c_double_p]
    lib.ns_forcing_box_extent.argtypes = [C.POINTER(orc.ForcingState), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.ns_forcing_active_indices.argtypes = [C.POINTER(orc.ForcingState), C.POINTER(C.c_int)]
    lib.ns_forcing_generate_box.restype = None
    lib.ns_forcing_generate_box.argtypes = [C.POINTER(orc.ForcingState), C.c_double, C.POINTER(C.c_int),
                                            C.c_int, C.c_int, c_double_p]
    lib.ns_config_init.restype = None
    lib.ns_config_init.argtypes = [C.POINTER(Config)]
    lib.ns_config_free.restype = None
    lib.ns_config_free.argtypes = [C.POINTER(Config)]
    lib.ns_config_set.argtypes = [C.POINTER(Config), C.c_char_p, C.c_char_p]
    lib.ns_config_get_string.restype = C.c_char_p
    lib.ns_config_get_string.argtypes = [C.POINTER(Config), C.c_char_p, C.c_char_p]
    lib.ns_config_get_int.argtypes = [C.POINTER(Config), C.c_char_p, C.c_int]
    lib.ns_config_get_double.restype = C.c_double
    lib.ns_config_get_double.argtypes = [C.POINTER(Config), C.c_char_p, C.c_double]
    lib.ns_config_get_bool.argtypes = [C.POINTER(Config), C.c_char_p, C.c_int]
    lib.ns_config_has_key.argtypes = [C.POINTER(Config), C.c_char_p]
    lib.ns_config_load_files.argtypes = [C.POINTER(Config), C.POINTER(C.c_char_p), C.c_size_t, C.c_char_p,
                                         C.c_size_t]
    lib.ns_config_validate.argtypes = [C.POINTER(Config), C.POINTER(SchemaEntry), C.c_size_t, C.c_char_p,
                                       C.c_size_t]
    return lib


class Solver:
    """A NSSpectralData instance driven exactly like the reference's tests drive it."""

    def __init__(self, lib, nx, ny, Lx=2 * np.pi, Ly=2 * np.pi, Re=100.0, plans=True):
        self.lib = lib
        self.sp = lib.ns_spectral_allocate(nx, ny, Lx, Ly, Re)
        assert self.sp
        lib.ns_spectral_initialize_grids(self.sp)
        if plans:
            lib.ns_spectral_initialize_fft_plans(self.sp)
            assert self.sp.contents.forward_plan, lib.ns_spectral_last_error()
        self.view = orc.View(self.sp)

    @property
    def s(self):
        return self.sp.contents

    def arr(self, name):
        return self.view.real(name) if name in orc.REAL_FIELDS else self.view.cplx(name)

    def close(self):
        if self.sp:
            self.lib.ns_spectral_free(self.sp)
            self.sp = None
