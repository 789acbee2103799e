/*
 * nsb_fft1024.cuh -- warp-level, register-resident FP64 FFT of length 1024.
 *
 * One warp owns one transform: lane t holds the 32 elements e = t + 32*m in registers
 * (x[m]). 1024 = 32 x 32:
 *   stage A   32-point DFT over m inside each thread (stride-32 elements)
 *   twiddle   x[r] *= w_1024^{t r}            (table TW[r*32 + t], shared by the CTA)
 *   transpose through shared memory, real and imaginary parts separately
 *             (32 x 33 doubles per warp, conflict-free both ways, __syncwarp only)
 *   stage B   32-point DFT over the former lane index
 * The result has the SAME distribution as the input (lane t holds X[t + 32*k'] in
 * x[k']), so pointwise work before/after the transform needs no permutation, and
 * global accesses are 512-byte coalesced per register index.
 *
 * Versus the shared-memory pass FFT in nsb_fft.cuh this does one smem round trip per
 * transform instead of four and no block-wide barrier at all (ncu, round 1: the pass
 * version is shared-memory bound with 45 % bank-conflict wavefronts).
 */
#ifndef NSB_FFT1024_CUH
#define NSB_FFT1024_CUH

#include "nsb_fft.cuh"

#define NSB_TP 33 /* pitch (doubles) of the per-warp transpose buffer */

__device__ __forceinline__ constexpr int nsb_bitrev5(int r) {
    return ((r & 1) << 4) | ((r & 2) << 2) | (r & 4) | ((r & 8) >> 2) | ((r & 16) >> 4);
}

/* o * exp(-2 pi i M/32); M is a compile-time constant after unrolling */
__device__ __forceinline__ cd nsb_twid32_c(cd o, int M) {
    constexpr double CS[16] = {1.0,
                               0.98078528040323044912618223613424,
                               0.92387953251128675612818318939679,
                               0.83146961230254523707878837761791,
                               0.70710678118654752440084436210485,
                               0.55557023301960222474283081394853,
                               0.38268343236508977172845998403040,
                               0.19509032201612826784828486847702,
                               0.0,
                               -0.19509032201612826784828486847702,
                               -0.38268343236508977172845998403040,
                               -0.55557023301960222474283081394853,
                               -0.70710678118654752440084436210485,
                               -0.83146961230254523707878837761791,
                               -0.92387953251128675612818318939679,
                               -0.98078528040323044912618223613424};
    constexpr double SN[16] = {0.0,
                               0.19509032201612826784828486847702,
                               0.38268343236508977172845998403040,
                               0.55557023301960222474283081394853,
                               0.70710678118654752440084436210485,
                               0.83146961230254523707878837761791,
                               0.92387953251128675612818318939679,
                               0.98078528040323044912618223613424,
                               1.0,
                               0.98078528040323044912618223613424,
                               0.92387953251128675612818318939679,
                               0.83146961230254523707878837761791,
                               0.70710678118654752440084436210485,
                               0.55557023301960222474283081394853,
                               0.38268343236508977172845998403040,
                               0.19509032201612826784828486847702};
    if (M == 0) return o;
    if (M == 8) return cmk(o.y, -o.x);
    if (M == 4) return cmk((o.x + o.y) * CS[4], (o.y - o.x) * CS[4]);
    if (M == 12) return cmk((o.y - o.x) * CS[4], -(o.x + o.y) * CS[4]);
    return cmk(o.x * CS[M] + o.y * SN[M], o.y * CS[M] - o.x * SN[M]);
}

/* forward 32-point DFT of x[0..32), natural order in and out, all in registers */
__device__ __forceinline__ void nsb_dft32_fwd(cd (&x)[32]) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
        for (int b = 0; b < 32; b += 2 * s) {
#pragma unroll
            for (int k = 0; k < s; ++k) {
                const cd a = x[b + k], c = x[b + k + s];
                x[b + k] = cadd(a, c);
                x[b + k + s] = nsb_twid32_c(csub(a, c), k * (16 / s));
            }
        }
    }
    cd y[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) y[r] = x[nsb_bitrev5(r)];
#pragma unroll
    for (int r = 0; r < 32; ++r) x[r] = y[r];
}

template <bool INV>
__device__ __forceinline__ void nsb_dft32(cd (&x)[32]) {
    if (INV) {
#pragma unroll
        for (int m = 0; m < 32; ++m) x[m] = cswap(x[m]);
    }
    nsb_dft32_fwd(x);
    if (INV) {
#pragma unroll
        for (int m = 0; m < 32; ++m) x[m] = cswap(x[m]);
    }
}

/* fills TW[r*32 + t] = exp(-2 pi i t r / 1024) from the per-context table TAB (the same 1024 values,
 * already in this order: a contiguous, coalesced 16 KB copy instead of a gather with stride r*t from the
 * length-n twiddle table, which every CTA of the column and comb kernels used to pay at start-up);
 * all threads of the CTA take part; caller syncs. */
__device__ __forceinline__ void nsb_fill_tw1024(cd *TW, const cd *__restrict__ TAB) {
    for (int k = threadIdx.x; k < 1024; k += blockDim.x) TW[k] = __ldg(&TAB[k]);
}

/*
 * In-place transform of the warp's 1024 elements. T: this warp's 32*NSB_TP doubles.
 * All 32 lanes must call. INV selects exp(+i).
 */
template <bool INV>
__device__ __forceinline__ void nsb_fft1024_warp(cd (&x)[32], double *T, const cd *TW, int lane) {
    nsb_dft32<INV>(x);
#pragma unroll
    for (int r = 1; r < 32; ++r) {
        const cd tw = TW[r * 32 + lane];
        x[r] = INV ? cmulc(x[r], tw) : cmul(x[r], tw);
    }
    /* transpose: lane q, register r  ->  lane r, register q */
#pragma unroll
    for (int r = 0; r < 32; ++r) T[r * NSB_TP + lane] = x[r].x;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 32; ++q) x[q].x = T[lane * NSB_TP + q];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; ++r) T[r * NSB_TP + lane] = x[r].y;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 32; ++q) x[q].y = T[lane * NSB_TP + q];
    __syncwarp();
    nsb_dft32<INV>(x);
}

/* warp FFT-1024 (nsb_fft1024.cuh) with an unpadded, XOR-swizzled 8 KB transpose buffer */
template <bool INV>
__device__ __forceinline__ void nsb_fft1024_warp_sw(cd (&x)[32], double *T, const cd *TW, int lane) {
    nsb_dft32<INV>(x);
#pragma unroll
    for (int r = 1; r < 32; ++r) {
        const cd tw = TW[r * 32 + lane];
        x[r] = INV ? cmulc(x[r], tw) : cmul(x[r], tw);
    }
#pragma unroll
    for (int r = 0; r < 32; ++r) T[r * 32 + (lane ^ r)] = x[r].x;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 32; ++q) x[q].x = T[lane * 32 + (q ^ lane)];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; ++r) T[r * 32 + (lane ^ r)] = x[r].y;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 32; ++q) x[q].y = T[lane * 32 + (q ^ lane)];
    __syncwarp();
    nsb_dft32<INV>(x);
}


#endif
