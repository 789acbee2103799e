/*
 * nsb_fft.cuh -- FP64 FFT building blocks for sm_100a.
 *
 *  - Dft<R>: radix-R (2..32) butterfly held entirely in registers, constant
 *    twiddles folded at compile time.
 *  - fft_smem_dif / fft_smem_dit: in-place mixed-radix passes over a tile of
 *    independent sequences staged in shared memory. DIF takes natural order in and
 *    leaves digit-reversed order; DIT takes digit-reversed in and leaves natural
 *    order. The kernels never un-permute: they address global memory through
 *    nsb_digitrev(), so the permutation costs nothing.
 *
 * These replace the FFTW3 calls of the reference
 * (/root/reference/src/spectral/spectral_transforms.c:10,31) together with the
 * 4-step splitting in nsb_kernels.cuh. Twiddles come from FP64 tables computed on
 * the host in long double (never sincos in the inner loop).
 */
#ifndef NSB_FFT_CUH
#define NSB_FFT_CUH

#include "nsb_platform.h"

__device__ __forceinline__ cd cmk(double a, double b) { return make_double2(a, b); }
__device__ __forceinline__ cd cadd(cd a, cd b) { return cmk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return cmk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd cmul(cd a, cd b) { return cmk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cd cmulc(cd a, cd b) { /* a * conj(b) */
    return cmk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ cd cconj(cd a) { return cmk(a.x, -a.y); }
__device__ __forceinline__ cd cswap(cd a) { return cmk(a.y, a.x); }
__device__ __forceinline__ cd cscale(cd a, double s) { return cmk(a.x * s, a.y * s); }

/* cos / sin of 2*pi*m/32 as compile-time constants */
NSB_HD constexpr double nsb_c32(int m) {
    return m == 0 ? 1.0
         : m == 1 ? 0.98078528040323044912618223613424
         : m == 2 ? 0.92387953251128675612818318939679
         : m == 3 ? 0.83146961230254523707878837761791
         : m == 4 ? 0.70710678118654752440084436210485
         : m == 5 ? 0.55557023301960222474283081394853
         : m == 6 ? 0.38268343236508977172845998403040
         : m == 7 ? 0.19509032201612826784828486847702
         : m == 8 ? 0.0
         : m < 16 ? -nsb_c32(16 - m)
         : m == 16 ? -1.0
                   : nsb_c32(32 - m);
}
NSB_HD constexpr double nsb_s32(int m) {
    return m <= 8 ? nsb_c32(8 - m) : m <= 16 ? nsb_s32(16 - m) : -nsb_s32(32 - m);
}

/* o * w, w = exp(-2*pi*i*M/32), M in [0,16) known at compile time */
template <int M>
__device__ __forceinline__ cd nsb_twid32(cd o) {
    if (M == 0) return o;
    if (M == 8) return cmk(o.y, -o.x);
    if (M == 4) {
        const double c = 0.70710678118654752440084436210485;
        return cmk((o.x + o.y) * c, (o.y - o.x) * c);
    }
    if (M == 12) {
        const double c = 0.70710678118654752440084436210485;
        return cmk((o.y - o.x) * c, -(o.x + o.y) * c);
    }
    const double C = nsb_c32(M), S = nsb_s32(M);
    return cmk(o.x * C + o.y * S, o.y * C - o.x * S);
}

template <int R> struct Dft;
template <> struct Dft<1> {
    __device__ static __forceinline__ void run(cd *) {}
};
template <> struct Dft<2> {
    __device__ static __forceinline__ void run(cd *x) {
        const cd a = x[0], b = x[1];
        x[0] = cadd(a, b);
        x[1] = csub(a, b);
    }
};
template <int R, int K> struct DftCombine {
    __device__ static __forceinline__ void run(const cd *e, const cd *o, cd *x) {
        const cd t = nsb_twid32<K * (32 / R)>(o[K]);
        x[K] = cadd(e[K], t);
        x[K + R / 2] = csub(e[K], t);
        DftCombine<R, K + 1>::run(e, o, x);
    }
};
template <> struct DftCombine<4, 2> { __device__ static __forceinline__ void run(const cd *, const cd *, cd *) {} };
template <> struct DftCombine<8, 4> { __device__ static __forceinline__ void run(const cd *, const cd *, cd *) {} };
template <> struct DftCombine<16, 8> { __device__ static __forceinline__ void run(const cd *, const cd *, cd *) {} };
template <> struct DftCombine<32, 16> { __device__ static __forceinline__ void run(const cd *, const cd *, cd *) {} };

/* forward DFT (e^{-i}) of x[0..R), natural order out */
template <int R> struct Dft {
    __device__ static __forceinline__ void run(cd *x) {
        cd e[R / 2], o[R / 2];
#pragma unroll
        for (int m = 0; m < R / 2; ++m) {
            e[m] = x[2 * m];
            o[m] = x[2 * m + 1];
        }
        Dft<R / 2>::run(e);
        Dft<R / 2>::run(o);
        DftCombine<R, 0>::run(e, o, x);
    }
};

template <int R, bool INV>
__device__ __forceinline__ void nsb_dft(cd *x) {
    if (INV) {
#pragma unroll
        for (int m = 0; m < R; ++m) x[m] = cswap(x[m]);
    }
    Dft<R>::run(x);
    if (INV) {
#pragma unroll
        for (int m = 0; m < R; ++m) x[m] = cswap(x[m]);
    }
}

/* radix used for a (sub)transform of length n; shared by DIF, DIT and digitrev */
NSB_HD inline int nsb_pick_radix(int n) { return (n % 8 == 0) ? 8 : (n % 4 == 0) ? 4 : 2; }

/* position p of a DIF output (== DIT input) holds index nsb_digitrev(p, N) */
NSB_HD inline int nsb_digitrev(int p, int N) {
    int idx = 0, mul = 1, n = N;
    while (n > 1) {
        const int R = nsb_pick_radix(n);
        n /= R;
        const int d = p / n;
        p -= d * n;
        idx += d * mul;
        mul *= R;
    }
    return idx;
}

/*
 * One radix-R pass over blocks of length n (n | N). Sequence b, position pos lives
 * at s[pos*SP + b*SB]. W[e*wmul] = exp(-2*pi*i*e/N) for e in [0,N).
 * DIT=false: butterfly then twiddle (decimation in frequency).
 * DIT=true : twiddle then butterfly (decimation in time).
 */
template <int R, bool INV, bool DIT>
__device__ __forceinline__ void nsb_fft_pass(cd *s, int N, int n, int nb, int SP, int SB,
                                             const cd *__restrict__ W, int wmul) {
    const int nR = n / R;
    const int nbf = N / R;           /* butterflies per sequence */
    const int items = nbf * nb;
    const int tstep = (N / n) * wmul; /* w_n^e = W[e*tstep] */
    for (int w = threadIdx.x; w < items; w += blockDim.x) {
        int b, g;
        if (SP == 1) { b = w / nbf; g = w - b * nbf; }  /* positions fastest across threads */
        else         { g = w / nb;  b = w - g * nb; }   /* sequences fastest across threads */
        const int blk = g / nR, q = g - blk * nR;
        cd *base = s + (size_t)(blk * n + q) * SP + (size_t)b * SB;
        const int estride = nR * SP;
        cd x[R];
#pragma unroll
        for (int m = 0; m < R; ++m) x[m] = base[m * estride];
        if (DIT) {
#pragma unroll
            for (int m = 1; m < R; ++m) {
                const cd tw = __ldg(&W[(size_t)(q * m) * tstep]);
                x[m] = INV ? cmulc(x[m], tw) : cmul(x[m], tw);
            }
        }
        nsb_dft<R, INV>(x);
        if (!DIT) {
#pragma unroll
            for (int m = 1; m < R; ++m) {
                const cd tw = __ldg(&W[(size_t)(q * m) * tstep]);
                x[m] = INV ? cmulc(x[m], tw) : cmul(x[m], tw);
            }
        }
#pragma unroll
        for (int m = 0; m < R; ++m) base[m * estride] = x[m];
    }
}

template <bool INV, bool DIT>
__device__ __forceinline__ void nsb_fft_pass_any(int R, cd *s, int N, int n, int nb, int SP, int SB,
                                                 const cd *__restrict__ W, int wmul) {
    if (R == 8) nsb_fft_pass<8, INV, DIT>(s, N, n, nb, SP, SB, W, wmul);
    else if (R == 4) nsb_fft_pass<4, INV, DIT>(s, N, n, nb, SP, SB, W, wmul);
    else nsb_fft_pass<2, INV, DIT>(s, N, n, nb, SP, SB, W, wmul);
}

/* natural in -> digit-reversed out. Ends with a __syncthreads(). Caller syncs before. */
template <bool INV>
__device__ __forceinline__ void nsb_fft_smem_dif(cd *s, int N, int nb, int SP, int SB,
                                                 const cd *__restrict__ W, int wmul) {
    int n = N;
    while (n > 1) {
        const int R = nsb_pick_radix(n);
        nsb_fft_pass_any<INV, false>(R, s, N, n, nb, SP, SB, W, wmul);
        __syncthreads();
        n /= R;
    }
}

/* digit-reversed in -> natural out. Ends with a __syncthreads(). */
template <bool INV>
__device__ __forceinline__ void nsb_fft_smem_dit(cd *s, int N, int nb, int SP, int SB,
                                                 const cd *__restrict__ W, int wmul) {
    int lev_n[16], lev_r[16], L = 0;
    int n = N;
    while (n > 1) {
        const int R = nsb_pick_radix(n);
        lev_n[L] = n; lev_r[L] = R; ++L;
        n /= R;
    }
    for (int l = L - 1; l >= 0; --l) {
        nsb_fft_pass_any<INV, true>(lev_r[l], s, N, lev_n[l], nb, SP, SB, W, wmul);
        __syncthreads();
    }
}

#endif
