/*
 * forcing.c -- band-limited stochastic forcing in omega_hat space (host side of the
 * hot path; the values are uploaded to the device as a small box, see
 * nsb_set_forcing in include/nsb.h).
 *
 * Stream contract of /root/reference/src/common/forcing.c, reproduced bit for bit
 * (checked against golden vectors minted from the unmodified reference,
 * tests/golden/forcing_*.json):
 *   splitmix64 (:11-16), uniform = (bits >> 11) * 2^-53 (:18-21), Box-Muller with
 *   u1 floored at DBL_MIN, cos branch returned first and the sin branch cached
 *   (:23-39); state seeded with cfg.seed, 0 -> 0x123456789abcdef (:99-100);
 *   ns_forcing_generate visits active modes in ascending index, draws two normals
 *   per mode, coloured: s <- alpha*s + beta*noise (:151-172), then rescales to
 *   energy_target (:176-183).
 * Build with -ffp-contract=off (as the oracle builds are) so the energy sum rounds
 * identically.
 */
#include "ns_forcing.h"

#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static uint64_t next_bits(unsigned long long *state) {
    uint64_t z = (*state += UINT64_C(0x9E3779B97F4A7C15));
    z = (z ^ (z >> 30)) * UINT64_C(0xBF58476D1CE4E5B9);
    z = (z ^ (z >> 27)) * UINT64_C(0x94D049BB133111EB);
    return z ^ (z >> 31);
}

static double next_uniform(NSForcingState *st) {
    return (double)(next_bits(&st->rng_state) >> 11) * (1.0 / 9007199254740992.0);
}

static double next_gaussian(NSForcingState *st) {
    if (st->have_spare) {
        st->have_spare = 0;
        return st->spare_gaussian;
    }
    double u1 = next_uniform(st);
    const double u2 = next_uniform(st);
    u1 = fmax(u1, DBL_MIN);
    const double radius = sqrt(-2.0 * log(u1));
    const double theta = 2.0 * M_PI * u2;
    st->spare_gaussian = radius * sin(theta);
    st->have_spare = 1;
    return radius * cos(theta);
}

void ns_forcing_config_default(NSForcingConfig *c) {
    if (!c) return;
    c->kmin = 0.0;
    c->kmax = 4.0;
    c->energy_target = 1.0;
    c->correlation_time = 0.5;
    c->seed = 0;
    c->type = NS_FORCING_WHITE_NOISE;
}

int ns_forcing_state_init(NSForcingState *st, const NSForcingConfig *c, const double *k2, int nkx, int nky) {
    if (!st || !c || !k2 || nkx <= 0 || nky <= 0) return -1;
    memset(st, 0, sizeof *st);
    st->config = *c;
    st->nkx = nkx;
    st->nky = nky;
    st->spectral_size = nkx * nky;
    const size_t m = (size_t)st->spectral_size;
    st->weights = (double *)calloc(m, sizeof(double));
    if (!st->weights) return -1;
    if (c->type == NS_FORCING_COLORED_NOISE) {
        st->colored_state = (double _Complex *)calloc(m, sizeof(double _Complex));
        if (!st->colored_state) {
            ns_forcing_state_free(st);
            return -1;
        }
    }
    const double lo = fmax(c->kmin, 0.0);
    const double hi = fmax(c->kmax, lo);
    int active = 0;
    for (size_t idx = 0; idx < m; ++idx) {
        const double kmag = sqrt(k2[idx]);
        if (kmag >= lo && kmag <= hi) {
            st->weights[idx] = 1.0;
            ++active;
        }
    }
    st->active_modes = active;
    st->have_spare = 0;
    st->rng_state = c->seed ? (unsigned long long)c->seed : 0x123456789abcdefULL;
    return active == 0 ? -1 : 0;
}

void ns_forcing_state_free(NSForcingState *st) {
    if (!st) return;
    free(st->weights);
    st->weights = NULL;
    free(st->colored_state);
    st->colored_state = NULL;
    st->spectral_size = 0;
    st->active_modes = 0;
}

/* alpha, beta of the Ornstein-Uhlenbeck update (forcing.c:138-147) */
static void ou_coefficients(const NSForcingConfig *c, double dt, double *alpha, double *beta) {
    *alpha = 0.0;
    *beta = 1.0;
    if (c->type == NS_FORCING_COLORED_NOISE) {
        const double tau = fmax(c->correlation_time, 1e-6);
        *alpha = exp(-dt / tau);
        *beta = sqrt(fmax(1.0 - (*alpha) * (*alpha), 0.0));
    }
}

/* one active mode: draw, colour, accumulate energy; returns the unscaled value */
static double _Complex draw_mode(NSForcingState *st, int idx, int colored, double alpha, double beta,
                                 double *energy) {
    const double g1 = next_gaussian(st);
    const double g2 = next_gaussian(st);
    const double _Complex noise = g1 + I * g2;
    double _Complex value = noise;
    if (colored) {
        st->colored_state[idx] = alpha * st->colored_state[idx] + beta * noise;
        value = st->colored_state[idx];
    }
    const double w = st->weights[idx];
    *energy += pow(w, 2.0) * (creal(value) * creal(value) + cimag(value) * cimag(value));
    return w * value;
}

void ns_forcing_generate(NSForcingState *st, double dt, const double *k2, double _Complex *out) {
    if (!st || !out || !k2) return;
    const int size = st->spectral_size;
    const size_t bytes = (size_t)size * sizeof(double _Complex);
    if (size == 0 || st->active_modes == 0) {
        memset(out, 0, bytes);
        return;
    }
    const NSForcingConfig *c = &st->config;
    if (c->energy_target <= 0.0) {
        memset(out, 0, bytes);
        if (st->colored_state) memset(st->colored_state, 0, bytes);
        return;
    }
    const int colored = (c->type == NS_FORCING_COLORED_NOISE);
    double alpha, beta;
    ou_coefficients(c, dt, &alpha, &beta);
    double energy = 0.0;
    for (int idx = 0; idx < size; ++idx) {
        if (st->weights[idx] <= 0.0) {
            out[idx] = 0.0;
            continue;
        }
        out[idx] = draw_mode(st, idx, colored, alpha, beta, &energy);
    }
    energy = fmax(energy, 1e-16);
    const double scale = sqrt(c->energy_target / energy);
    for (int idx = 0; idx < size; ++idx) {
        out[idx] *= scale;
        if (colored) st->colored_state[idx] *= scale;
    }
}

/* ------------------------------------------------------------ additive: box path */

int ns_forcing_active_indices(const NSForcingState *st, int *indices) {
    if (!st || !indices || !st->weights) return -1;
    int n = 0;
    for (int idx = 0; idx < st->spectral_size; ++idx)
        if (st->weights[idx] > 0.0) indices[n++] = idx;
    return n == st->active_modes ? 0 : -1;
}

int ns_forcing_box_extent(const NSForcingState *st, int *ib, int *jb) {
    if (!st || !ib || !jb || !st->weights || st->active_modes <= 0) return -1;
    int mi = 0, mj = 0;
    for (int idx = 0; idx < st->spectral_size; ++idx) {
        if (st->weights[idx] <= 0.0) continue;
        const int j = idx / st->nkx, i = idx - j * st->nkx;
        const int band = (j <= st->nky / 2) ? j : j - st->nky;
        if (i > mi) mi = i;
        if (abs(band) > mj) mj = abs(band);
    }
    *ib = mi;
    *jb = mj;
    return 0;
}

void ns_forcing_generate_box(NSForcingState *st, double dt, const int *indices, int ib, int jb,
                             double _Complex *box) {
    if (!st || !indices || !box) return;
    const size_t cells = (size_t)(2 * jb + 1) * (size_t)(ib + 1);
    memset(box, 0, cells * sizeof(double _Complex));
    if (st->spectral_size == 0 || st->active_modes == 0) return;
    const NSForcingConfig *c = &st->config;
    const int colored = (c->type == NS_FORCING_COLORED_NOISE);
    if (c->energy_target <= 0.0) {
        if (st->colored_state)
            for (int a = 0; a < st->active_modes; ++a) st->colored_state[indices[a]] = 0.0;
        return;
    }
    double alpha, beta;
    ou_coefficients(c, dt, &alpha, &beta);
    double energy = 0.0;
    for (int a = 0; a < st->active_modes; ++a) {
        const int idx = indices[a];
        const int j = idx / st->nkx, i = idx - j * st->nkx;
        const int band = (j <= st->nky / 2) ? j : j - st->nky;
        box[(size_t)(band + jb) * (size_t)(ib + 1) + (size_t)i] = draw_mode(st, idx, colored, alpha, beta, &energy);
    }
    energy = fmax(energy, 1e-16);
    const double scale = sqrt(c->energy_target / energy);
    for (size_t k = 0; k < cells; ++k) box[k] *= scale;
    if (colored)
        for (int a = 0; a < st->active_modes; ++a) st->colored_state[indices[a]] *= scale;
}
