/*
 * nsb_stats.cuh -- turbulence statistics of a 2-D velocity field on the device.
 *
 * Replaces the host loops of /root/reference/src/analysis/statistics.c:
 *   :189-201  means                          k_stats_sums
 *   :219-239  central moments 2..4           k_stats_moments (same pass also does the two below)
 *   :104-130  Taylor microscale  <(du'/dx)^2> with centred periodic differences
 *   :132-175  Kolmogorov scale   <Sxx^2 + Syy^2 + 2 Sxy^2>
 *   :10-75    two-point correlations of u' along x and y, direct periodic sums  k_stats_corr
 * The scalar post-processing (rms, skewness, kurtosis, integral scales by the trapezoid rule with the
 * first-zero-crossing cut, :77-102, :241-262, :318-327) is a handful of flops and stays on the host
 * (csrc/host/analysis.c) with the reference's operation order.
 * Any nx, ny >= 1 (no transform involved). Sums are block-reduced and combined with double atomics, so
 * the summation order varies run to run at the 1e-16 level, like the reference's OpenMP reductions.
 */
#ifndef NSB_STATS_CUH
#define NSB_STATS_CUH

#include "nsb_kernels.cuh"

/* all threads call; red: blockDim doubles of shared memory */
template <int NV>
__device__ __forceinline__ void nsb_block_sums_to_global(const double (&v)[NV], double *dst, double *red) {
#pragma unroll
    for (int k = 0; k < NV; ++k) nsb_block_sum_to_global(v[k], dst + k, red);
}

__global__ void __launch_bounds__(256) k_stats_sums(const double *u, const double *v, size_t n, double *out) {
    NSB_DYN_SMEM(double, red);
    double s[2] = {0.0, 0.0};
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        s[0] += u[k];
        s[1] += v[k];
    }
    nsb_block_sums_to_global<2>(s, out, red);
}

/* out[0..2] = sum du^2, du^3, du^4; out[3..5] the same for v; out[6] = sum (du'/dx)^2; out[7] = sum of the
 * strain invariant */
__global__ void __launch_bounds__(256) k_stats_moments(const double *u, const double *v, int nx, int ny, double mean_u,
                                                       double mean_v, double dx, double dy, double *out) {
    NSB_DYN_SMEM(double, red);
    const size_t n = (size_t)nx * ny;
    double s[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(k / nx), j = (int)(k - (size_t)i * nx);
        const size_t row = (size_t)i * nx;
        const size_t jp = row + (j + 1 == nx ? 0 : j + 1), jm = row + (j == 0 ? nx - 1 : j - 1);
        const size_t ip = (size_t)(i + 1 == ny ? 0 : i + 1) * nx + j, im = (size_t)(i == 0 ? ny - 1 : i - 1) * nx + j;
        const double du = u[k] - mean_u, dv = v[k] - mean_v;
        const double du2 = du * du, dv2 = dv * dv;
        s[0] += du2; s[1] += du2 * du; s[2] += du2 * du2;
        s[3] += dv2; s[4] += dv2 * dv; s[5] += dv2 * dv2;
        /* statistics.c:121 on the fluctuation field */
        const double dfx = ((u[jp] - mean_u) - (u[jm] - mean_u)) / (2.0 * dx);
        s[6] += dfx * dfx;
        /* statistics.c:155-164 on the full fields */
        const double du_dx = (u[jp] - u[jm]) / (2.0 * dx), du_dy = (u[ip] - u[im]) / (2.0 * dy);
        const double dv_dx = (v[jp] - v[jm]) / (2.0 * dx), dv_dy = (v[ip] - v[im]) / (2.0 * dy);
        const double Sxy = 0.5 * (du_dy + dv_dx);
        s[7] += du_dx * du_dx + dv_dy * dv_dy + 2.0 * Sxy * Sxy;
    }
    nsb_block_sums_to_global<8>(s, out, red);
}

/* corr[s] += sum over a chunk of points of u'(p) u'(p shifted by s along x (dir 0) or y (dir 1)), periodic.
 * grid (chunks, nsep): the shifted operand of a chunk comes from the same rows (x) or from rows s below (y), so
 * after the first separation the field is served by L2. */
__global__ void __launch_bounds__(256) k_stats_corr(const double *u, int nx, int ny, double mean_u, int dir,
                                                    double *corr) {
    NSB_DYN_SMEM(double, red);
    const size_t n = (size_t)nx * ny;
    const int s = blockIdx.y;
    double acc = 0.0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(k / nx), j = (int)(k - (size_t)i * nx);
        size_t q;
        if (dir == 0) {
            int jj = j + s;
            if (jj >= nx) jj -= nx;
            q = (size_t)i * nx + jj;
        } else {
            int ii = i + s;
            if (ii >= ny) ii -= ny;
            q = (size_t)ii * nx + j;
        }
        acc += (u[k] - mean_u) * (u[q] - mean_u);
    }
    nsb_block_sum_to_global(acc, corr + s, red);
}

#endif
