/*
 * oracle/ns_oracle.h -- TEST INFRASTRUCTURE ONLY (parity checker + CPU baseline).
 *
 * CPU restatement ("port") of the reference's Fourier pseudospectral time step,
 * /root/reference/src/spectral/{spectral_core,spectral_time_integration,
 * spectral_transforms,hyperviscosity,spectral_initialization,spectral_analysis}.c
 * and /root/reference/src/common/forcing.c. Every function cites the reference
 * file:line it follows. The FFT is oracle/fft_cpu.c (FFTW3 is an un-vendored
 * system dependency of the reference, version unpinned; 3.3.10 per
 * docs/benchmark_results.md:72, absent from this image).
 *
 * PINNING: this restatement is checked (tests/test_oracle.py) against
 *   (a) oracle/_ref/libns_ref.so = the UNMODIFIED reference sources compiled
 *       here against oracle/shim/fftw3.h, whenever /root/reference is present;
 *   (b) the golden vectors in tests/golden/ which were produced by (a) with
 *       oracle/gen_golden.py;
 *   (c) the assertions of the reference's own passing tests
 *       (tests/test_solvers.c:239-295,433-664; tests/test_conservation.c:62-186).
 *
 * Nothing outside tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * The state struct mirrors NSSpectralData field for field
 * (src/spectral/ns_spectral_solver.h:22-68) so one ctypes description serves the
 * reference build, this port and the product.
 */
#ifndef NS_ORACLE_H
#define NS_ORACLE_H

#include <complex.h>

#ifdef __cplusplus
extern "C" {
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
} OrcHyperviscosity;

typedef struct {
    int nx, ny;
    int nkx, nky;
    double Lx, Ly;
    double Re;
    double dt;
    double t;
    double cfl_factor;
    double *u, *v, *p, *omega, *psi, *x, *y;
    double complex *u_hat, *v_hat, *omega_hat, *psi_hat, *nonlinear_hat;
    double complex *rk1_hat, *rk2_hat, *rk3_hat, *rk4_hat, *omega_tmp_hat;
    double *kx, *ky, *k2;
    void *forward_plan, *backward_plan;
    int *dealias_mask;
    OrcHyperviscosity hyperviscosity;
    double *hyperviscosity_weights;
    double complex *sfd_hat;
} OrcSpectral;

typedef struct {
    double kmin, kmax, energy_target, correlation_time;
    unsigned long seed;
    int type; /* 0 white, 1 coloured */
} OrcForcingConfig;

typedef struct {
    OrcForcingConfig config;
    int nkx, nky, spectral_size, active_modes;
    double *weights;
    double complex *colored_state;
    unsigned int have_spare;
    double spare_gaussian;
    unsigned long long rng_state;
} OrcForcingState;

OrcSpectral *orc_spectral_allocate(int nx, int ny, double Lx, double Ly, double Re);
void orc_spectral_free(OrcSpectral *s);
void orc_spectral_initialize_grids(OrcSpectral *s);
void orc_spectral_initialize_fft_plans(OrcSpectral *s);
void orc_spectral_initialize_taylor_green(OrcSpectral *s);
void orc_spectral_physical_to_spectral(OrcSpectral *s, double *field, double complex *field_hat);
void orc_spectral_spectral_to_physical(OrcSpectral *s, double complex *field_hat, double *field);
void orc_spectral_apply_dealiasing(OrcSpectral *s, double complex *field_hat);
void orc_spectral_compute_derivatives(OrcSpectral *s);
void orc_spectral_compute_velocity_from_vorticity(OrcSpectral *s);
void orc_spectral_compute_nonlinear_terms(OrcSpectral *s);
void orc_spectral_advance_rk4(OrcSpectral *s);
double orc_spectral_compute_cfl_timestep(OrcSpectral *s);
double orc_spectral_compute_energy(OrcSpectral *s);
double orc_spectral_compute_enstrophy(OrcSpectral *s);
void orc_spectral_hyperviscosity_default(OrcHyperviscosity *c);
void orc_spectral_hyperviscosity_set(OrcSpectral *s, const OrcHyperviscosity *c);
void orc_spectral_hyperviscosity_update_weights(OrcSpectral *s);
void orc_spectral_hyperviscosity_apply(OrcSpectral *s, double complex *rhs_hat, const double complex *state_hat);
void orc_spectral_hyperviscosity_post_step(OrcSpectral *s, const double complex *state_hat, double dt);

void orc_forcing_config_default(OrcForcingConfig *c);
int orc_forcing_state_init(OrcForcingState *st, const OrcForcingConfig *c, const double *k2, int nkx, int nky);
void orc_forcing_state_free(OrcForcingState *st);
void orc_forcing_generate(OrcForcingState *st, double dt, const double *k2, double complex *out);

/* Additions the reference lacks; defined identically in
 * oracle/ref_driver.c on top of the reference's public functions. */
void orc_forced_rk4(OrcSpectral *s, const double complex *forcing /* nullable */);
void orc_random_phase_init(OrcSpectral *s, unsigned long long seed, double k0, double max_omega);
/* raw generator outputs for known-answer tests */
void orc_rng_kat(unsigned long long seed, int n, unsigned long long *bits, double *uniforms, double *gaussians);

#ifdef __cplusplus
}
#endif
#endif
