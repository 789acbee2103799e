/*
 * include/ns_forcing.h -- stochastic forcing generator (host side of the hot path).
 *
 * Drop-in for /root/reference/src/common/forcing.h:7-46: same types, same four
 * entry points. The RNG stream (splitmix64 -> 53-bit uniforms -> Box-Muller, cos
 * branch first, sin branch cached; forcing.c:11-39) is reproduced bit for bit, and
 * ns_forcing_generate consumes it in the reference's order (ascending mode index).
 *
 * Additive: ns_forcing_generate_box() is the O(active modes) path the device loop
 * uses; it advances the state exactly like ns_forcing_generate() and returns the
 * same values packed into the bounding box of the active band.
 */
#ifndef NS_COMMON_FORCING_H
#define NS_COMMON_FORCING_H

#include "common_types.h"
#include <complex.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    NS_FORCING_WHITE_NOISE = 0,
    NS_FORCING_COLORED_NOISE = 1
} NSForcingType;

typedef struct {
    double kmin;
    double kmax;
    double energy_target;
    double correlation_time;
    unsigned long seed;
    NSForcingType type;
} NSForcingConfig;

typedef struct {
    NSForcingConfig config;
    int nkx;
    int nky;
    int spectral_size;
    int active_modes;
    double *weights;
    double _Complex *colored_state;
    unsigned int have_spare;
    double spare_gaussian;
    unsigned long long rng_state;
} NSForcingState;

void ns_forcing_config_default(NSForcingConfig *config);
int ns_forcing_state_init(NSForcingState *state, const NSForcingConfig *config, const double *k2,
                          int nkx, int nky);
void ns_forcing_state_free(NSForcingState *state);
void ns_forcing_generate(NSForcingState *state, double dt, const double *k2, double _Complex *forcing_out);

/* ---- additive ---- */
/* Bounding box of the active band: i in [0,*ib], |ky band| <= *jb. Returns 0 / -1. */
int ns_forcing_box_extent(const NSForcingState *state, int *ib, int *jb);
/* Same stream consumption and values as ns_forcing_generate, but only the active modes
 * are visited (list supplied by the caller: ascending indices from
 * ns_forcing_active_indices) and the result lands in box[(band+jb)*(ib+1)+i]. */
int ns_forcing_active_indices(const NSForcingState *state, int *indices /* active_modes entries */);
void ns_forcing_generate_box(NSForcingState *state, double dt, const int *indices, int ib, int jb,
                             double _Complex *box);

#ifdef __cplusplus
}
#endif
#endif /* NS_COMMON_FORCING_H */
