/*
 * oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiled INTO oracle/_ref/libns_ref.so next to the UNMODIFIED reference
 * sources (see oracle/Makefile). Adds the two things the reference lacks
 *  using only the reference's PUBLIC functions:
 *
 *   nsref_forced_rk4        ns_spectral_advance_rk4 verbatim in structure
 *                           (/root/reference/src/spectral/spectral_time_integration.c:34-106)
 *                           with "+ f_hat" added to every stage RHS, f_hat held
 *                           constant over the step (paper/manuscript.tex:60-63).
 *                           With forcing == NULL it must reproduce
 *                           ns_spectral_advance_rk4 bit for bit (tested).
 *   nsref_random_phase_init build-defined turbulence initial condition (8d).
 *
 * Same definitions as orc_forced_rk4 / orc_random_phase_init in ns_oracle.c.
 */
#include "ns_spectral_solver.h" /* the reference's own header, via -I */
#include <stdint.h>
#include <string.h>

static void rhs_into(NSSpectralData *d, fftw_complex *k, const fftw_complex *forcing) {
    const int m = d->nkx * d->nky;
    const double inv_Re = 1.0 / d->Re;
    ns_spectral_compute_nonlinear_terms(d);
    for (int q = 0; q < m; ++q)
        k[q] = d->nonlinear_hat[q] - d->k2[q] * d->omega_hat[q] * inv_Re;
    ns_spectral_hyperviscosity_apply(d, k, d->omega_hat);
    if (forcing)
        for (int q = 0; q < m; ++q) k[q] += forcing[q];
}

static void advance_to(NSSpectralData *d, double c, const fftw_complex *k) {
    const int m = d->nkx * d->nky;
    for (int q = 0; q < m; ++q) d->omega_hat[q] = d->omega_tmp_hat[q] + c * k[q];
    ns_spectral_compute_derivatives(d);
}

void nsref_forced_rk4(NSSpectralData *d, const fftw_complex *forcing) {
    const int m = d->nkx * d->nky;
    const double dt = d->dt;
    memcpy(d->omega_tmp_hat, d->omega_hat, (size_t)m * sizeof(fftw_complex));
    rhs_into(d, d->rk1_hat, forcing);
    advance_to(d, 0.5 * dt, d->rk1_hat);
    rhs_into(d, d->rk2_hat, forcing);
    advance_to(d, 0.5 * dt, d->rk2_hat);
    rhs_into(d, d->rk3_hat, forcing);
    advance_to(d, dt, d->rk3_hat);
    rhs_into(d, d->rk4_hat, forcing);
    for (int q = 0; q < m; ++q)
        d->omega_hat[q] = d->omega_tmp_hat[q] +
                          dt / 6.0 * (d->rk1_hat[q] + 2.0 * d->rk2_hat[q] + 2.0 * d->rk3_hat[q] + d->rk4_hat[q]);
    ns_spectral_hyperviscosity_post_step(d, d->omega_hat, dt);
    d->t += dt;
    ns_spectral_compute_derivatives(d);
}

static uint64_t sm64(uint64_t *st) {
    uint64_t z = (*st += UINT64_C(0x9E3779B97F4A7C15));
    z = (z ^ (z >> 30)) * UINT64_C(0xBF58476D1CE4E5B9);
    z = (z ^ (z >> 27)) * UINT64_C(0x94D049BB133111EB);
    return z ^ (z >> 31);
}

void nsref_random_phase_init(NSSpectralData *d, unsigned long long seed, double k0, double max_omega) {
    const int nkx = d->nkx, nky = d->nky, ny = d->ny;
    const int n = d->nx * d->ny;
    uint64_t st = seed;
    for (int j = 0; j < nky; ++j) {
        const int band = (j <= ny / 2) ? j : j - ny;
        for (int i = 0; i < nkx; ++i) {
            const double un = (double)(sm64(&st) >> 11) * (1.0 / 9007199254740992.0);
            const double theta = 2.0 * M_PI * un;
            const double kk = sqrt((double)i * i + (double)band * band);
            const double a = kk * kk * exp(-(kk / k0) * (kk / k0));
            d->omega_hat[j * nkx + i] = a * cos(theta) + I * (a * sin(theta));
        }
    }
    d->omega_hat[0] = 0.0;
    ns_spectral_spectral_to_physical(d, d->omega_hat, d->omega);
    double peak = 0.0;
    for (int q = 0; q < n; ++q) peak = fmax(peak, fabs(d->omega[q]));
    const double scale = (peak > 0.0) ? max_omega / peak : 0.0;
    for (int q = 0; q < n; ++q) d->omega[q] *= scale;
    ns_spectral_physical_to_spectral(d, d->omega, d->omega_hat);
    ns_spectral_compute_derivatives(d);
}

/* The reference allocates with fftw_alloc_* (not zeroed). Parity runs
 * call this right after ns_spectral_allocate so both sides start from zeros. */
void nsref_zero_fill(NSSpectralData *d) {
    const size_t n = (size_t)d->nx * d->ny, m = (size_t)d->nkx * d->nky;
    double *r[] = {d->u, d->v, d->p, d->omega, d->psi};
    fftw_complex *c[] = {d->u_hat, d->v_hat, d->omega_hat, d->psi_hat, d->nonlinear_hat,
                         d->rk1_hat, d->rk2_hat, d->rk3_hat, d->rk4_hat, d->omega_tmp_hat};
    for (size_t a = 0; a < sizeof(r) / sizeof(r[0]); ++a) memset(r[a], 0, n * sizeof(double));
    for (size_t a = 0; a < sizeof(c) / sizeof(c[0]); ++a) memset(c[a], 0, m * sizeof(fftw_complex));
}
