/*
 * include/ns_statistics.h -- turbulence statistics of a 2-D velocity field, B200 backend.
 *
 * Drop-in for /root/reference/src/analysis/statistics.h: same two structs (field order kept) and the
 * same entry point with the same argument meaning and return convention (:38-45 there); the sums over
 * the grid run on the device (navier-stokes-solvers_b200/csrc/cuda/nsb_stats.cuh), the scalar
 * post-processing on the host in the reference's operation order. No CPU fallback: without a CUDA
 * device the call returns -1 (ns_spectral_last_error() has the text).
 */
#ifndef NS_STATISTICS_H
#define NS_STATISTICS_H

#include "common_types.h"
#include "ns_spectral_solver.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Central moments of a scalar field (statistics.h:11-16) */
typedef struct {
    double mean;
    double rms;
    double skewness;
    double kurtosis;
} NSFieldMoments;

/* Collection of turbulence statistics for a 2D velocity field (statistics.h:19-26) */
typedef struct {
    NSFieldMoments u;
    NSFieldMoments v;
    double integral_length_scale_x;
    double integral_length_scale_y;
    double taylor_microscale;
    double kolmogorov_scale;
} NSTurbulenceStatistics;

/* statistics.h:38-45 / statistics.c:177-339. corr_x / corr_y (optional) receive the normalised two-point
 * correlations of the u fluctuation along x / y, max_sep_x / max_sep_y entries. Any nx, ny. Returns 0 on
 * success, non-zero on invalid input, allocation failure or a missing device. The device used is the one
 * NS_SPECTRAL_DEVICE names (default 0). */
int ns_stats_compute_velocity_statistics(const double *NS_RESTRICT u, const double *NS_RESTRICT v, int nx, int ny,
                                         double dx, double dy, double nu, int max_sep_x, int max_sep_y,
                                         NSTurbulenceStatistics *stats, double *corr_x, double *corr_y);

/* Additive: the same statistics of the SOLVER's current velocity field with dx = Lx/nx, dy = Ly/ny,
 * nu = 1/Re. Under NS_SYNC_LAZY u and v are formed from the device-resident omega_hat and never leave
 * the device; under NS_SYNC_EAGER the host arrays data->u, data->v are used like the reference would. */
int ns_spectral_turbulence_statistics(NSSpectralData *data, int max_sep_x, int max_sep_y,
                                      NSTurbulenceStatistics *stats, double *corr_x, double *corr_y);

#ifdef __cplusplus
}
#endif
#endif
