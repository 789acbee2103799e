/*
 * include/ns_spectral_solver.h -- the spectral solver's C entry points, B200 backend.
 *
 * Drop-in for /root/reference/src/spectral/ns_spectral_solver.h:
 *   - NSSpectralHyperviscosityConfig (:8-19) and NSSpectralData (:22-68) keep every
 *     field, in the same order with the same names; ONE field is appended
 *     (`void *backend`, opaque). Callers that poke the arrays directly
 *     (tests/test_solvers.c:521-526,636-660) keep working: see the coherence policy.
 *   - the 22 prototypes (:71-113) keep their signatures.
 *   - <fftw3.h> is NOT included: fftw_complex is C99 `double _Complex`, fftw_plan an
 *     opaque pointer (FFTW is replaced by the sm_100a Stockham kernels behind
 *     include/nsb.h).
 *
 * Coherence policy (NSSpectralData is host-visible, the solver state lives in HBM):
 *   NS_SYNC_EAGER (default)  every entry point uploads the struct arrays it reads and
 *       downloads everything the reference would have written. Reference tests pass
 *       unchanged, at PCIe cost.
 *   NS_SYNC_LAZY             omega_hat (and sfd_hat) live on the device; entry points
 *       move no field data; the host arrays are refreshed only by
 *       ns_spectral_sync_to_host(). The CLI and every benchmark run this way.
 *   Mid-call clobbering of u,v,psi,omega,p by compute_nonlinear_terms
 *   (spectral_core.c:54-59) is not reproduced; only end-of-call state is contractual.
 *
 * There is no CPU fallback: without a visible CUDA device
 * ns_spectral_initialize_fft_plans() reports the error on stderr, leaves the plans
 * NULL, and every later numerical entry point aborts loudly.
 */
#ifndef NS_SPECTRAL_SOLVER_H
#define NS_SPECTRAL_SOLVER_H

#include "common_types.h"
#include "ns_forcing.h"
#include <complex.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef NS_HAVE_FFTW_TYPES
#define NS_HAVE_FFTW_TYPES
typedef double _Complex fftw_complex;
typedef void *fftw_plan;
#endif

typedef struct {
    int enabled;
    double coefficient;
    int order;
    double cutoff_ratio;
    double sfd_strength;
    double sfd_timescale;
    int sfd_enabled;
    double kmax;
    double cutoff_wavenumber;
    double inv_bandwidth;
} NSSpectralHyperviscosityConfig;

typedef struct {
    int nx, ny;       /* grid dimensions */
    int nkx, nky;     /* Fourier mode dimensions (nx/2+1, ny) */
    double Lx, Ly;    /* domain size */
    double Re;        /* Reynolds number */
    double dt;        /* time step */
    double t;         /* current time */
    double cfl_factor;

    double *NS_RESTRICT u;
    double *NS_RESTRICT v;
    double *NS_RESTRICT p;
    double *NS_RESTRICT omega;
    double *NS_RESTRICT psi;
    double *NS_RESTRICT x;
    double *NS_RESTRICT y;

    fftw_complex *NS_RESTRICT u_hat;
    fftw_complex *NS_RESTRICT v_hat;
    fftw_complex *NS_RESTRICT omega_hat;
    fftw_complex *NS_RESTRICT psi_hat;
    fftw_complex *NS_RESTRICT nonlinear_hat;
    fftw_complex *NS_RESTRICT rk1_hat;
    fftw_complex *NS_RESTRICT rk2_hat;
    fftw_complex *NS_RESTRICT rk3_hat;
    fftw_complex *NS_RESTRICT rk4_hat;
    fftw_complex *NS_RESTRICT omega_tmp_hat;

    double *NS_RESTRICT kx;
    double *NS_RESTRICT ky;
    double *NS_RESTRICT k2;

    fftw_plan forward_plan, backward_plan; /* non-NULL once the device context exists */

    int *NS_RESTRICT dealias_mask;

    NSSpectralHyperviscosityConfig hyperviscosity;
    double *NS_RESTRICT hyperviscosity_weights;
    fftw_complex *NS_RESTRICT sfd_hat;

    void *backend; /* appended: opaque device backend (include/nsb.h context + policy) */
} NSSpectralData;

/* ---- the reference's 22 entry points (ns_spectral_solver.h:71-113) ---- */
NSSpectralData *ns_spectral_allocate(int nx, int ny, double Lx, double Ly, double Re);
void ns_spectral_free(NSSpectralData *data);

void ns_spectral_initialize_grids(NSSpectralData *data);
void ns_spectral_initialize_fft_plans(NSSpectralData *data);
void ns_spectral_initialize_taylor_green(NSSpectralData *data);

void ns_spectral_physical_to_spectral(NSSpectralData *data, double *field, fftw_complex *field_hat);
void ns_spectral_spectral_to_physical(NSSpectralData *data, fftw_complex *field_hat, double *field);
void ns_spectral_apply_dealiasing(NSSpectralData *data, fftw_complex *field_hat);

void ns_spectral_compute_derivatives(NSSpectralData *data);
void ns_spectral_compute_velocity_from_vorticity(NSSpectralData *data);
void ns_spectral_compute_nonlinear_terms(NSSpectralData *data);

void ns_spectral_advance_rk4(NSSpectralData *data);
double ns_spectral_compute_cfl_timestep(NSSpectralData *data);

double ns_spectral_compute_energy(NSSpectralData *data);
double ns_spectral_compute_enstrophy(NSSpectralData *data);
void ns_spectral_analyze_convergence(NSSpectralData *data, int step);
void ns_spectral_analyze_spectrum(NSSpectralData *data);
int ns_spectral_check_resolution(NSSpectralData *data);

void ns_spectral_hyperviscosity_default(NSSpectralHyperviscosityConfig *config);
void ns_spectral_hyperviscosity_set(NSSpectralData *data, const NSSpectralHyperviscosityConfig *config);
void ns_spectral_hyperviscosity_update_weights(NSSpectralData *data);
void ns_spectral_hyperviscosity_apply(NSSpectralData *data, fftw_complex *rhs_hat, const fftw_complex *state_hat);
void ns_spectral_hyperviscosity_post_step(NSSpectralData *data, const fftw_complex *state_hat, double dt);

void ns_spectral_output_solution(NSSpectralData *data, const char *filename);

/* ---- additive API (absent from the reference) ---- */
typedef enum { NS_SYNC_EAGER = 0, NS_SYNC_LAZY = 1 } NSSyncPolicy;

/* field masks for the explicit syncs */
#define NS_FIELD_OMEGA_HAT (1u << 0)
#define NS_FIELD_PSI_HAT   (1u << 1)
#define NS_FIELD_U_HAT     (1u << 2)
#define NS_FIELD_V_HAT     (1u << 3)
#define NS_FIELD_SFD_HAT   (1u << 4)
#define NS_FIELD_U         (1u << 8)
#define NS_FIELD_V         (1u << 9)
#define NS_FIELD_OMEGA     (1u << 10)
#define NS_FIELD_PSI       (1u << 11)
#define NS_FIELD_DERIVED   (NS_FIELD_PSI_HAT | NS_FIELD_U_HAT | NS_FIELD_V_HAT | NS_FIELD_U | NS_FIELD_V | NS_FIELD_OMEGA)
#define NS_FIELD_ALL       (NS_FIELD_OMEGA_HAT | NS_FIELD_SFD_HAT | NS_FIELD_DERIVED | NS_FIELD_PSI)

/* CUDA device used by the next ns_spectral_initialize_fft_plans (default 0 or
 * $NS_SPECTRAL_DEVICE). Returns 0, or -1 if the context already exists. */
int ns_spectral_backend_select_device(NSSpectralData *data, int device);
/* slab-decompose over n_gpus devices driven by this process (after initialize_fft_plans);
 * ns_spectral_run_steps then advances the one grid on all of them. 0 / -1 + ns_spectral_last_error() */
int ns_spectral_backend_init(NSSpectralData *data, int n_gpus);
void ns_spectral_set_sync_policy(NSSpectralData *data, NSSyncPolicy policy);
NSSyncPolicy ns_spectral_get_sync_policy(const NSSpectralData *data);
/* device -> host: derived fields are recomputed on the device from omega_hat first */
int ns_spectral_sync_to_host(NSSpectralData *data, unsigned field_mask);
/* host -> device (omega_hat, sfd_hat are the only device-owned state) */
int ns_spectral_sync_to_device(NSSpectralData *data, unsigned field_mask);

/* ns_spectral_advance_rk4 with "+ forcing" in every stage RHS, forcing constant over
 * the step (paper/manuscript.tex:60-63). forcing == NULL is advance_rk4. */
int ns_spectral_advance_rk4_forced(NSSpectralData *data, const fftw_complex *forcing);

/* The driver loop `dt = cfl(); advance_rk4()` (main_spectral.c:46-48) run n times
 * entirely on the device: adaptive dt from the fused max|u| reduction, no field
 * transfers. `forcing` (nullable) is advanced exactly like ns_forcing_generate would,
 * one call per step. On return data->t and data->dt hold the device values; under
 * NS_SYNC_EAGER the host arrays are refreshed as after advance_rk4. */
int ns_spectral_run_steps(NSSpectralData *data, int n, NSForcingState *forcing);

/* Build-defined turbulence initial condition: seeded random phases with
 * amplitude kk^2 exp(-(kk/k0)^2), rescaled to max|omega| = max_omega. */
int ns_spectral_initialize_random_phase(NSSpectralData *data, unsigned long long seed, double k0,
                                        double max_omega);

/* ns_spectral_output_solution with every `stride`-th point in x and y (stride 1 = the reference's
 * file; spectral_output.c:4-27). 0 / -1. */
int ns_spectral_output_solution_strided(NSSpectralData *data, const char *filename, int stride);

/* Shell sums behind ns_spectral_analyze_spectrum (spectral_analysis.c:103-142): ek_sum[b] = sum of
 * 0.5 (|u_hat|^2 + |v_hat|^2), count[b] = modes with b = (int)(sqrt(k2) + 0.5) < nbins. Under
 * NS_SYNC_LAZY computed on the device from omega_hat, no download. 0 / -1. */
int ns_spectral_energy_spectrum(NSSpectralData *data, int nbins, double *ek_sum, int *count);

/* Restartable state: omega_hat, the SFD filter state, t, dt and (forcing nullable) the forcing
 * generator's stream position and coloured-noise state in one raw binary file; a run resumed with
 * ns_spectral_load_state continues bit for bit like the uninterrupted one. The solver (and the
 * forcing state) must have been set up for the same grid (band, type) before loading. The
 * reference's HDF5 checkpoint (src/common/checkpoint.c:266-301) omits the filter and forcing
 * state. 0 / -1 + ns_spectral_last_error() */
int ns_spectral_save_state(NSSpectralData *data, const NSForcingState *forcing, const char *filename);
int ns_spectral_load_state(NSSpectralData *data, NSForcingState *forcing, const char *filename);

/* last error text of the calling thread ("" if none) */
const char *ns_spectral_last_error(void);
/* the nsb context behind `data` (include/nsb.h), NULL before initialize_fft_plans */
void *ns_spectral_backend_context(NSSpectralData *data);

#ifdef __cplusplus
}
#endif
#endif /* NS_SPECTRAL_SOLVER_H */
