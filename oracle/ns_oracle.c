/*
 * oracle/ns_oracle.c -- TEST INFRASTRUCTURE ONLY. See ns_oracle.h.
 *
 * CPU restatement of the reference spectral time step. The per-element
 * arithmetic (operand order included) follows the cited reference lines; the
 * code organisation is this repo's own (one RHS helper instead of four
 * unrolled stages, one derivative helper instead of four loops, ...).
 */
#define _DEFAULT_SOURCE
#include "ns_oracle.h"
#include "fft_cpu.h"

#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define ORC_MIN_DT 1e-8 /* include/common_types.h:32 */
#define ORC_MAX_DT 0.1  /* include/common_types.h:33 */

typedef double complex cplx;

static void *xalloc(size_t bytes) {
    void *p = NULL;
    if (posix_memalign(&p, 64, bytes ? bytes : 64) != 0) return NULL;
    return p;
}

/* ------------------------------------------------------------------ memory */

/* spectral_memory.c:6-68 -- sizes, defaults (dt 0.01, cfl 0.3), zeroed weights
 * and sfd_hat. Unlike the reference every array is zero-filled (SURVEY G7:
 * psi_hat is read before it is first written). */
OrcSpectral *orc_spectral_allocate(int nx, int ny, double Lx, double Ly, double Re) {
    OrcSpectral *s = (OrcSpectral *)calloc(1, sizeof(*s));
    if (!s) return NULL;
    s->nx = nx; s->ny = ny;
    s->nkx = nx / 2 + 1; s->nky = ny;
    s->Lx = Lx; s->Ly = Ly; s->Re = Re;
    s->dt = 0.01; s->t = 0.0; s->cfl_factor = 0.3;
    const size_t n = (size_t)nx * ny, m = (size_t)s->nkx * s->nky;
    double **reals[] = {&s->u, &s->v, &s->p, &s->omega, &s->psi};
    cplx **cplxs[] = {&s->u_hat, &s->v_hat, &s->omega_hat, &s->psi_hat, &s->nonlinear_hat,
                      &s->rk1_hat, &s->rk2_hat, &s->rk3_hat, &s->rk4_hat, &s->omega_tmp_hat,
                      &s->sfd_hat};
    int ok = 1;
    for (size_t a = 0; a < sizeof(reals) / sizeof(reals[0]); ++a) {
        *reals[a] = (double *)xalloc(n * sizeof(double));
        if (*reals[a]) memset(*reals[a], 0, n * sizeof(double)); else ok = 0;
    }
    for (size_t a = 0; a < sizeof(cplxs) / sizeof(cplxs[0]); ++a) {
        *cplxs[a] = (cplx *)xalloc(m * sizeof(cplx));
        if (*cplxs[a]) memset(*cplxs[a], 0, m * sizeof(cplx)); else ok = 0;
    }
    s->x = (double *)calloc((size_t)nx, sizeof(double));
    s->y = (double *)calloc((size_t)ny, sizeof(double));
    s->kx = (double *)calloc((size_t)s->nkx, sizeof(double));
    s->ky = (double *)calloc((size_t)s->nky, sizeof(double));
    s->k2 = (double *)calloc(m, sizeof(double));
    s->dealias_mask = (int *)calloc(m, sizeof(int));
    s->hyperviscosity_weights = (double *)calloc(m, sizeof(double));
    if (!ok || !s->x || !s->y || !s->kx || !s->ky || !s->k2 || !s->dealias_mask ||
        !s->hyperviscosity_weights) {
        orc_spectral_free(s);
        return NULL;
    }
    orc_spectral_hyperviscosity_default(&s->hyperviscosity);
    return s;
}

/* spectral_memory.c:71-104 */
void orc_spectral_free(OrcSpectral *s) {
    if (!s) return;
    fft2d_plan_destroy((fft2d_plan *)s->forward_plan);
    free(s->u); free(s->v); free(s->p); free(s->omega); free(s->psi);
    free(s->u_hat); free(s->v_hat); free(s->omega_hat); free(s->psi_hat);
    free(s->nonlinear_hat); free(s->rk1_hat); free(s->rk2_hat); free(s->rk3_hat);
    free(s->rk4_hat); free(s->omega_tmp_hat); free(s->sfd_hat);
    free(s->x); free(s->y); free(s->kx); free(s->ky); free(s->k2);
    free(s->dealias_mask); free(s->hyperviscosity_weights);
    free(s);
}

/* ---------------------------------------------------------- initialisation */

/* spectral_initialization.c:10-54 */
void orc_spectral_initialize_grids(OrcSpectral *s) {
    const int nx = s->nx, ny = s->ny, nkx = s->nkx, nky = s->nky;
    for (int i = 0; i < nx; ++i) s->x[i] = i * s->Lx / nx;                 /* :16 */
    for (int j = 0; j < ny; ++j) s->y[j] = j * s->Ly / ny;                 /* :20 */
    for (int i = 0; i < nkx; ++i) s->kx[i] = 2.0 * M_PI * i / s->Lx;       /* :26 */
    for (int j = 0; j < nky; ++j) {                                        /* :31-35 */
        const int band = (j <= ny / 2) ? j : j - ny;
        s->ky[j] = 2.0 * M_PI * band / s->Ly;
    }
    const int xcut = (2 * nkx) / 3;   /* :46 integer division on nkx, not nx */
    const int ycut = (2 * ny) / 6;    /* :48 */
    for (int j = 0; j < nky; ++j) {
        const int band = (j <= ny / 2) ? j : j - ny;
        const int keep_y = abs(band) <= ycut;
        for (int i = 0; i < nkx; ++i) {
            const size_t idx = (size_t)j * nkx + i;
            s->k2[idx] = s->kx[i] * s->kx[i] + s->ky[j] * s->ky[j];        /* :44 */
            s->dealias_mask[idx] = (i <= xcut) && keep_y;                  /* :49 */
        }
    }
    orc_spectral_hyperviscosity_update_weights(s);                         /* :53 */
}

/* spectral_initialization.c:57-69 (one plan object serves both directions) */
void orc_spectral_initialize_fft_plans(OrcSpectral *s) {
    if (s->forward_plan) return;
    s->forward_plan = fft2d_plan_create(s->ny, s->nx);
    s->backward_plan = s->forward_plan;
}

/* spectral_initialization.c:72-95 (G8: omega is NOT curl of this u,v; kept) */
void orc_spectral_initialize_taylor_green(OrcSpectral *s) {
    const int nx = s->nx, ny = s->ny;
    const double Lx = s->Lx, Ly = s->Ly;
#pragma omp parallel for
    for (int j = 0; j < ny; ++j) {
        for (int i = 0; i < nx; ++i) {
            const size_t idx = (size_t)j * nx + i;
            const double x = s->x[i], y = s->y[j];
            s->u[idx] = sin(2 * M_PI * x / Lx) * cos(2 * M_PI * y / Ly);
            s->v[idx] = -cos(2 * M_PI * x / Lx) * sin(2 * M_PI * y / Ly);
            s->omega[idx] = 2.0 * (2 * M_PI / Lx) * cos(2 * M_PI * x / Lx) * cos(2 * M_PI * y / Ly);
            s->p[idx] = 0.0;
        }
    }
    orc_spectral_physical_to_spectral(s, s->omega, s->omega_hat);
    orc_spectral_physical_to_spectral(s, s->u, s->u_hat);
    orc_spectral_physical_to_spectral(s, s->v, s->v_hat);
}

/* -------------------------------------------------------------- transforms */

/* spectral_transforms.c:35-43 */
void orc_spectral_apply_dealiasing(OrcSpectral *s, cplx *f) {
    const size_t m = (size_t)s->nkx * s->nky;
#pragma omp parallel for
    for (size_t k = 0; k < m; ++k)
        if (!s->dealias_mask[k]) f[k] = 0.0;
}

/* spectral_transforms.c:9-22 : r2c, times 1/(nx*ny), mask */
void orc_spectral_physical_to_spectral(OrcSpectral *s, double *field, cplx *field_hat) {
    fft2d_r2c((const fft2d_plan *)s->forward_plan, field, field_hat);
    const double norm = 1.0 / (s->nx * s->ny);
    const size_t m = (size_t)s->nkx * s->nky;
#pragma omp parallel for
    for (size_t k = 0; k < m; ++k) field_hat[k] *= norm;
    orc_spectral_apply_dealiasing(s, field_hat);
}

/* spectral_transforms.c:25-32 : copy to scratch (= nonlinear_hat), mask, c2r */
void orc_spectral_spectral_to_physical(OrcSpectral *s, cplx *field_hat, double *field) {
    const size_t m = (size_t)s->nkx * s->nky;
    cplx *scratch = s->nonlinear_hat;
    if (scratch != field_hat) memcpy(scratch, field_hat, m * sizeof(cplx));
    orc_spectral_apply_dealiasing(s, scratch);
    fft2d_c2r((const fft2d_plan *)s->backward_plan, scratch, field);
}

/* -------------------------------------------------------------------- core */

/* spectral_core.c:8-34 : psi_hat = -omega_hat/k2 (0 where k2 <= 1e-12),
 * u_hat = i ky psi_hat, v_hat = -i kx psi_hat */
void orc_spectral_compute_velocity_from_vorticity(OrcSpectral *s) {
    const int nkx = s->nkx, nky = s->nky;
#pragma omp parallel for
    for (int j = 0; j < nky; ++j) {
        for (int i = 0; i < nkx; ++i) {
            const size_t idx = (size_t)j * nkx + i;
            if (s->k2[idx] > 1e-12) s->psi_hat[idx] = -s->omega_hat[idx] / s->k2[idx];
            else s->psi_hat[idx] = 0.0;
            s->u_hat[idx] = I * s->ky[j] * s->psi_hat[idx];
            s->v_hat[idx] = -I * s->kx[i] * s->psi_hat[idx];
        }
    }
}

/* spectral_core.c:37-43 */
void orc_spectral_compute_derivatives(OrcSpectral *s) {
    orc_spectral_compute_velocity_from_vorticity(s);
    orc_spectral_spectral_to_physical(s, s->u_hat, s->u);
    orc_spectral_spectral_to_physical(s, s->v_hat, s->v);
    orc_spectral_spectral_to_physical(s, s->omega_hat, s->omega);
}

/* one of the four "i k f_hat -> scratch -> c2r" blocks, spectral_core.c:61-99 */
static void ik_derivative(OrcSpectral *s, const cplx *f_hat, int along_y, double *out) {
    const int nkx = s->nkx, nky = s->nky;
    cplx *scratch = s->nonlinear_hat;
#pragma omp parallel for
    for (int j = 0; j < nky; ++j) {
        for (int i = 0; i < nkx; ++i) {
            const size_t idx = (size_t)j * nkx + i;
            const double k = along_y ? s->ky[j] : s->kx[i];
            scratch[idx] = I * k * f_hat[idx];
        }
    }
    orc_spectral_spectral_to_physical(s, scratch, out);
}

/* spectral_core.c:46-112 : N_hat = -mask*FFT[psi_x*omega_y - psi_y*omega_x]/(nx ny).
 * The reference's two leading c2r of psi_hat and omega_hat (:51-52) only produce
 * arrays that are overwritten at :54-57 and are skipped here. The
 * aliasing of the physical arrays (:54-59) is reproduced so end-of-call state
 * matches: psi_x->u, psi_y->v, omega_x->psi, omega_y->omega, J->p. */
void orc_spectral_compute_nonlinear_terms(OrcSpectral *s) {
    const size_t n = (size_t)s->nx * s->ny, m = (size_t)s->nkx * s->nky;
    double *psi_x = s->u, *psi_y = s->v, *omega_x = s->psi, *omega_y = s->omega, *jac = s->p;
    ik_derivative(s, s->psi_hat, 0, psi_x);
    ik_derivative(s, s->psi_hat, 1, psi_y);
    ik_derivative(s, s->omega_hat, 0, omega_x);
    ik_derivative(s, s->omega_hat, 1, omega_y);
#pragma omp parallel for
    for (size_t k = 0; k < n; ++k) jac[k] = psi_x[k] * omega_y[k] - psi_y[k] * omega_x[k];   /* :103 */
    orc_spectral_physical_to_spectral(s, jac, s->nonlinear_hat);
#pragma omp parallel for
    for (size_t k = 0; k < m; ++k) s->nonlinear_hat[k] = -s->nonlinear_hat[k];               /* :110 */
}

/* --------------------------------------------------------- time integration */

/* spectral_time_integration.c:9-31 */
double orc_spectral_compute_cfl_timestep(OrcSpectral *s) {
    const size_t n = (size_t)s->nx * s->ny;
    double max_vel = 0.0;
#pragma omp parallel for reduction(max : max_vel)
    for (size_t k = 0; k < n; ++k) {
        const double mag = sqrt(s->u[k] * s->u[k] + s->v[k] * s->v[k]);
        if (mag > max_vel) max_vel = mag;
    }
    const double dx_min = fmin(s->Lx / s->nx, s->Ly / s->ny);
    const double dt_cfl = s->cfl_factor * dx_min / (max_vel + 1e-12);
    const double dt_visc = 0.5 * s->Re * dx_min * dx_min;
    double dt = fmin(dt_cfl, dt_visc);
    dt = fmax(dt, ORC_MIN_DT);
    dt = fmin(dt, ORC_MAX_DT);
    return dt;
}

/* RHS of one stage: spectral_time_integration.c:49-54 (and :62-67, :75-80, :88-93),
 * plus the forcing term of SURVEY 8c when `forcing` is given. */
static void stage_rhs(OrcSpectral *s, cplx *k_out, const cplx *forcing) {
    const size_t m = (size_t)s->nkx * s->nky;
    const double inv_Re = 1.0 / s->Re;
    orc_spectral_compute_nonlinear_terms(s);
#pragma omp parallel for
    for (size_t k = 0; k < m; ++k)
        k_out[k] = s->nonlinear_hat[k] - s->k2[k] * s->omega_hat[k] * inv_Re;
    orc_spectral_hyperviscosity_apply(s, k_out, s->omega_hat);
    if (forcing) {
#pragma omp parallel for
        for (size_t k = 0; k < m; ++k) k_out[k] += forcing[k];
    }
}

/* omega_hat = omega_0 + c*k, then refresh psi_hat etc.: :57-61, :70-74, :83-87 */
static void stage_advance(OrcSpectral *s, double c, const cplx *k_in) {
    const size_t m = (size_t)s->nkx * s->nky;
    const cplx *w0 = s->omega_tmp_hat;
#pragma omp parallel for
    for (size_t k = 0; k < m; ++k) s->omega_hat[k] = w0[k] + c * k_in[k];
    orc_spectral_compute_derivatives(s);
}

/* spectral_time_integration.c:34-106 with an optional constant-in-step forcing
 * (paper/manuscript.tex:60-63). forcing == NULL is advance_rk4. */
void orc_forced_rk4(OrcSpectral *s, const cplx *forcing) {
    const size_t m = (size_t)s->nkx * s->nky;
    const double dt = s->dt;
    cplx *k1 = s->rk1_hat, *k2 = s->rk2_hat, *k3 = s->rk3_hat, *k4 = s->rk4_hat;
    cplx *w0 = s->omega_tmp_hat;
    memcpy(w0, s->omega_hat, m * sizeof(cplx));           /* :44 */
    stage_rhs(s, k1, forcing);                             /* stage 1 uses the stored psi_hat */
    stage_advance(s, 0.5 * dt, k1);
    stage_rhs(s, k2, forcing);
    stage_advance(s, 0.5 * dt, k2);
    stage_rhs(s, k3, forcing);
    stage_advance(s, dt, k3);
    stage_rhs(s, k4, forcing);
#pragma omp parallel for
    for (size_t k = 0; k < m; ++k)                          /* :97-100 */
        s->omega_hat[k] = w0[k] + dt / 6.0 * (k1[k] + 2.0 * k2[k] + 2.0 * k3[k] + k4[k]);
    orc_spectral_hyperviscosity_post_step(s, s->omega_hat, dt);   /* :102 */
    s->t += dt;                                                    /* :104 */
    orc_spectral_compute_derivatives(s);                           /* :105 */
}

void orc_spectral_advance_rk4(OrcSpectral *s) { orc_forced_rk4(s, NULL); }

/* ------------------------------------------------------------- diagnostics */

/* spectral_analysis.c:8-19 */
double orc_spectral_compute_energy(OrcSpectral *s) {
    const size_t n = (size_t)s->nx * s->ny;
    double e = 0.0;
#pragma omp parallel for reduction(+ : e)
    for (size_t k = 0; k < n; ++k) e += 0.5 * (s->u[k] * s->u[k] + s->v[k] * s->v[k]);
    return e / (s->nx * s->ny);
}

/* spectral_analysis.c:22-33 */
double orc_spectral_compute_enstrophy(OrcSpectral *s) {
    const size_t n = (size_t)s->nx * s->ny;
    double z = 0.0;
#pragma omp parallel for reduction(+ : z)
    for (size_t k = 0; k < n; ++k) z += 0.5 * s->omega[k] * s->omega[k];
    return z / (s->nx * s->ny);
}

/* ---------------------------------------------------------- hyperviscosity */

static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* hyperviscosity.c:15-30 */
void orc_spectral_hyperviscosity_default(OrcHyperviscosity *c) {
    if (!c) return;
    memset(c, 0, sizeof(*c));
    c->order = 8;
    c->cutoff_ratio = 0.65;
    c->sfd_timescale = 1.0;
}

/* hyperviscosity.c:32-58 */
void orc_spectral_hyperviscosity_set(OrcSpectral *s, const OrcHyperviscosity *c) {
    if (!s || !c) return;
    const size_t m = (size_t)s->nkx * s->nky;
    s->hyperviscosity = *c;
    OrcHyperviscosity *h = &s->hyperviscosity;
    if (!h->enabled) memset(s->hyperviscosity_weights, 0, m * sizeof(double));
    if (!h->sfd_enabled || h->sfd_strength <= 0.0) {
        h->sfd_enabled = 0;
        h->sfd_strength = 0.0;
        memset(s->sfd_hat, 0, m * sizeof(cplx));
    }
    orc_spectral_hyperviscosity_update_weights(s);
    if (h->sfd_enabled) memcpy(s->sfd_hat, s->omega_hat, m * sizeof(cplx));
}

/* hyperviscosity.c:60-102 : kmax over ALL modes; w = c*((|k|-kc)/(kmax-kc))^order */
void orc_spectral_hyperviscosity_update_weights(OrcSpectral *s) {
    if (!s || !s->hyperviscosity_weights || !s->k2) return;
    const size_t m = (size_t)s->nkx * s->nky;
    OrcHyperviscosity *h = &s->hyperviscosity;
    double kmax = 0.0;
    for (size_t k = 0; k < m; ++k) kmax = fmax(kmax, sqrt(s->k2[k]));
    h->kmax = kmax;
    h->cutoff_wavenumber = clampd(h->cutoff_ratio, 0.0, 1.0) * kmax;
    const double bw = fmax(kmax - h->cutoff_wavenumber, 1e-12);
    h->inv_bandwidth = (bw > 0.0) ? 1.0 / bw : 0.0;
    if (!h->enabled || h->coefficient <= 0.0 || kmax <= 0.0) {
        memset(s->hyperviscosity_weights, 0, m * sizeof(double));
        return;
    }
    const int order = (h->order > 0) ? h->order : 1;
#pragma omp parallel for
    for (size_t k = 0; k < m; ++k) {
        const double kmag = sqrt(s->k2[k]);
        double w = 0.0;
        if (kmag > h->cutoff_wavenumber) {
            const double ratio = (kmag - h->cutoff_wavenumber) * h->inv_bandwidth;
            w = h->coefficient * pow(ratio, order);
        }
        s->hyperviscosity_weights[k] = w;
    }
}

/* hyperviscosity.c:104-150 */
void orc_spectral_hyperviscosity_apply(OrcSpectral *s, cplx *rhs, const cplx *state) {
    if (!s || !rhs || !state) return;
    const OrcHyperviscosity *h = &s->hyperviscosity;
    const int visc = (h->enabled && h->coefficient > 0.0);
    const int sfd = (h->sfd_enabled && h->sfd_strength > 0.0);
    if (!visc && !sfd) return;
    const int order = (h->order > 0) ? h->order : 1;
    const size_t m = (size_t)s->nkx * s->nky;
#pragma omp parallel for
    for (size_t k = 0; k < m; ++k) {
        cplx damp = 0.0;
        if (visc) damp += s->hyperviscosity_weights[k] * state[k];
        if (sfd) {
            double ratio = 0.0;
            if (h->inv_bandwidth > 0.0) {
                const double kmag = sqrt(s->k2[k]);
                if (kmag > h->cutoff_wavenumber) {
                    const double sc = clampd((kmag - h->cutoff_wavenumber) * h->inv_bandwidth, 0.0, 1.0);
                    ratio = pow(sc, order);
                }
            }
            const double sw = h->sfd_strength * ratio;
            if (sw > 0.0) damp += sw * (state[k] - s->sfd_hat[k]);
        }
        rhs[k] -= damp;
    }
}

/* hyperviscosity.c:152-178 */
void orc_spectral_hyperviscosity_post_step(OrcSpectral *s, const cplx *state, double dt) {
    if (!s || !state || dt <= 0.0) return;
    const OrcHyperviscosity *h = &s->hyperviscosity;
    if (!h->sfd_enabled || h->sfd_strength <= 0.0 || h->sfd_timescale <= 0.0) return;
    const double alpha = clampd(dt / h->sfd_timescale, 0.0, 1.0);
    const size_t m = (size_t)s->nkx * s->nky;
#pragma omp parallel for
    for (size_t k = 0; k < m; ++k)
        s->sfd_hat[k] = (1.0 - alpha) * s->sfd_hat[k] + alpha * state[k];
}

/* ----------------------------------------------------------------- forcing */

/* forcing.c:11-16 */
static uint64_t splitmix64_next(uint64_t *st) {
    uint64_t z = (*st += UINT64_C(0x9E3779B97F4A7C15));
    z = (z ^ (z >> 30)) * UINT64_C(0xBF58476D1CE4E5B9);
    z = (z ^ (z >> 27)) * UINT64_C(0x94D049BB133111EB);
    return z ^ (z >> 31);
}

/* forcing.c:18-21 : top 53 bits times 2^-53 */
static double uniform01(uint64_t *st) {
    return (double)(splitmix64_next(st) >> 11) * (1.0 / 9007199254740992.0);
}

/* forcing.c:23-39 : Box-Muller, cos branch returned first, sin branch cached */
static double gaussian(OrcForcingState *st) {
    if (st->have_spare) {
        st->have_spare = 0;
        return st->spare_gaussian;
    }
    uint64_t rs = st->rng_state;
    double u1 = uniform01(&rs);
    const double u2 = uniform01(&rs);
    st->rng_state = rs;
    u1 = fmax(u1, DBL_MIN);
    const double radius = sqrt(-2.0 * log(u1));
    const double theta = 2.0 * M_PI * u2;
    st->spare_gaussian = radius * sin(theta);
    st->have_spare = 1;
    return radius * cos(theta);
}

/* forcing.c:41-51 */
void orc_forcing_config_default(OrcForcingConfig *c) {
    if (!c) return;
    c->kmin = 0.0; c->kmax = 4.0; c->energy_target = 1.0;
    c->correlation_time = 0.5; c->seed = 0; c->type = 0;
}

/* forcing.c:53-103 */
int orc_forcing_state_init(OrcForcingState *st, const OrcForcingConfig *c, const double *k2,
                           int nkx, int nky) {
    if (!st || !c || !k2 || nkx <= 0 || nky <= 0) return -1;
    memset(st, 0, sizeof(*st));
    st->config = *c;
    st->nkx = nkx; st->nky = nky; st->spectral_size = nkx * nky;
    st->weights = (double *)calloc((size_t)st->spectral_size, sizeof(double));
    if (!st->weights) return -1;
    if (c->type == 1) {
        st->colored_state = (cplx *)calloc((size_t)st->spectral_size, sizeof(cplx));
        if (!st->colored_state) { orc_forcing_state_free(st); return -1; }
    }
    const double kmin = fmax(c->kmin, 0.0);
    const double kmax = fmax(c->kmax, kmin);
    int active = 0;
    for (int idx = 0; idx < st->spectral_size; ++idx) {
        const double kmag = sqrt(k2[idx]);
        if (kmag >= kmin && kmag <= kmax) { st->weights[idx] = 1.0; ++active; }
    }
    st->active_modes = active;
    st->rng_state = c->seed ? (uint64_t)c->seed : UINT64_C(0x123456789abcdef);
    return active == 0 ? -1 : 0;
}

/* forcing.c:105-115 */
void orc_forcing_state_free(OrcForcingState *st) {
    if (!st) return;
    free(st->weights); st->weights = NULL;
    free(st->colored_state); st->colored_state = NULL;
    st->spectral_size = 0; st->active_modes = 0;
}

/* forcing.c:117-184 */
void orc_forcing_generate(OrcForcingState *st, double dt, const double *k2, cplx *out) {
    if (!st || !out || !k2) return;
    const int size = st->spectral_size;
    if (size == 0 || st->active_modes == 0) { memset(out, 0, (size_t)size * sizeof(cplx)); return; }
    const OrcForcingConfig *c = &st->config;
    if (c->energy_target <= 0.0) {
        memset(out, 0, (size_t)size * sizeof(cplx));
        if (st->colored_state) memset(st->colored_state, 0, (size_t)size * sizeof(cplx));
        return;
    }
    const double tau = fmax(c->correlation_time, 1e-6);
    const int colored = (c->type == 1);
    double alpha = 0.0, beta = 1.0;
    if (colored) {
        alpha = exp(-dt / tau);
        beta = sqrt(fmax(1.0 - alpha * alpha, 0.0));
    }
    double total = 0.0;
    for (int idx = 0; idx < size; ++idx) {
        if (st->weights[idx] <= 0.0) { out[idx] = 0.0; continue; }
        const double g1 = gaussian(st);
        const double g2 = gaussian(st);
        const cplx noise = g1 + I * g2;
        cplx value = noise;
        if (colored) {
            st->colored_state[idx] = alpha * st->colored_state[idx] + beta * noise;
            value = st->colored_state[idx];
        }
        out[idx] = st->weights[idx] * value;
        total += pow(st->weights[idx], 2.0) * (creal(value) * creal(value) + cimag(value) * cimag(value));
    }
    total = fmax(total, 1e-16);
    const double scale = sqrt(c->energy_target / total);
    for (int idx = 0; idx < size; ++idx) {
        out[idx] *= scale;
        if (colored) st->colored_state[idx] *= scale;
    }
}

void orc_rng_kat(unsigned long long seed, int n, unsigned long long *bits, double *uniforms,
                 double *gaussians) {
    uint64_t s1 = seed, s2 = seed;
    for (int k = 0; k < n; ++k) {
        if (bits) bits[k] = splitmix64_next(&s1);
        if (uniforms) uniforms[k] = uniform01(&s2);
    }
    if (gaussians) {
        OrcForcingState st;
        memset(&st, 0, sizeof(st));
        st.rng_state = seed;
        for (int k = 0; k < n; ++k) gaussians[k] = gaussian(&st);
    }
}

/* ------------------------------------------------- random-phase initial data */

/* Build-defined (the reference has no turbulence IC); SURVEY 8d. One uniform per
 * mode in ascending idx over ALL modes, omega_hat = a(k) e^{i theta},
 * a = kk^2 exp(-(kk/k0)^2) with kk the index-space |k|, (0,0) zeroed; then
 * c2r -> rescale to max|omega| = max_omega -> physical_to_spectral ->
 * compute_derivatives. */
void orc_random_phase_init(OrcSpectral *s, unsigned long long seed, double k0, double max_omega) {
    const int nkx = s->nkx, nky = s->nky, ny = s->ny;
    const size_t n = (size_t)s->nx * ny;
    uint64_t st = seed;
    for (int j = 0; j < nky; ++j) {
        const int band = (j <= ny / 2) ? j : j - ny;
        for (int i = 0; i < nkx; ++i) {
            const double theta = 2.0 * M_PI * uniform01(&st);
            const double kk = sqrt((double)i * i + (double)band * band);
            const double a = kk * kk * exp(-(kk / k0) * (kk / k0));
            s->omega_hat[(size_t)j * nkx + i] = a * cos(theta) + I * (a * sin(theta));
        }
    }
    s->omega_hat[0] = 0.0;
    orc_spectral_spectral_to_physical(s, s->omega_hat, s->omega);
    double peak = 0.0;
    for (size_t k = 0; k < n; ++k) peak = fmax(peak, fabs(s->omega[k]));
    const double scale = (peak > 0.0) ? max_omega / peak : 0.0;
    for (size_t k = 0; k < n; ++k) s->omega[k] *= scale;
    orc_spectral_physical_to_spectral(s, s->omega, s->omega_hat);
    orc_spectral_compute_derivatives(s);
}
