/*
 * nsb_kernels.cuh -- the sm_100a kernels of the pseudospectral RK4 step.
 *
 * Replaces, for the hot path of /root/reference/src/spectral:
 *   spectral_core.c:8-34,46-112       (Poisson solve, ik-derivatives, Jacobian)
 *   spectral_transforms.c:9-43        (r2c/c2r + 1/(nx ny) + 2/3 mask)
 *   spectral_time_integration.c:9-106 (CFL dt, RK4 stages, final combine)
 *   hyperviscosity.c:104-178          (explicit damping, SFD filter)
 *
 * 2-D transforms are split as ny = Y1*Y2, nx = X1*X2 ("4-step" in both
 * directions) so that every on-chip tile fits one SM at any grid size:
 *
 *   alpha_inv : spectral pointwise (psi = -w/k^2, i k derivatives, mask) + inverse
 *               length-Y1 column FFTs + y twiddle            -> T[f][y1][j2][col]
 *   beta_inv  : length-Y2 DFT over j2, Hermitian gather, length-X1 DFT over k1,
 *               x twiddle                                     -> H[f][y][x1][k2]
 *   gamma     : two packed complex length-X2 inverse FFTs per comb (psi_x+i psi_y,
 *               w_x+i w_y), Jacobian + max|u|^2 in registers/smem, one packed
 *               forward length-X2 FFT for two combs          -> H'[y][x1][k2]
 *   beta_fwd  : length-X1 DFT over x1, length-Y2 DFT over y2  -> T'[y1][j2][col]
 *   alpha_fwd : y twiddle + forward length-Y1 column FFTs + (-1/(nx ny), mask) +
 *               RHS assembly, damping, forcing, RK stage update in the epilogue.
 *
 * When Y2 == X1 == 1 the beta kernels vanish and a stage is three kernels.
 * Masked columns (kx index > (2 nkx)/3) are never transformed or stored; masked
 * rows are never loaded. Physical-space fields never touch HBM in the step.
 */
#ifndef NSB_KERNELS_CUH
#define NSB_KERNELS_CUH

#include "nsb_fft.cuh"
#include "nsb_fft1024.cuh"

#define NSB_MAXP 8 /* devices of one NVSwitch domain */
#define NSB_TWK_PITCH(X1) ((X1) + ((X1) < 32 ? 1 : 0)) /* row pitch (complex elements) of the batch-ordered x twiddle table */

struct NsbGeom {
    int nx, ny, nkx;
    int Y1, Y2, X1, X2, X2h;
    int imax, ncol, ycut;
    int PT, PH;      /* row pitches (complex elements) of the T and H layouts */
    int C;           /* columns per alpha CTA */
    int BC;          /* k2 values per beta CTA */
    size_t fsT, fsH; /* field strides (complex elements) */
    const double *kx, *ky;
    const double *kyK; /* ky permuted for the fast column kernel: kyK[j2*Y1 + j1] = ky[j1*Y2 + j2] */
    const cd *WY, *WX; /* exp(-2 pi i m/ny), exp(-2 pi i m/nx) */
    const cd *TWY, *TWX; /* warp FFT-1024 twiddles exp(-2 pi i t r/1024) at [r*32 + t] (ny, nx >= 1024 only) */
    const cd *TWK;       /* x twiddles of the chain's B kernel in batch order: TWK[k2*(X1+1) + x1] = WX[k2*x1], k2 <= X2h
                          * (padded with 1 to a multiple of 16 k2): one contiguous block per k2 batch */
    double norm;       /* 1/(nx*ny) */
};

struct NsbHv {
    int visc, sfd, order;
    double coef, kc, inv_bw, sfd_strength, sfd_timescale;
};

struct NsbRk {
    int stage;
    cd *W0, *Wst, *ACC, *SFD;
    const double *dt;
    double inv_Re;
    NsbHv hv;
    const cd *FB; /* forcing box [(band+jb)*(ib+1)+i], nullable */
    int ib, jb;
};

/* max that keeps a NaN once it has seen one (fmax drops NaNs): a blown-up field must not turn into
 * "max |u| = 0, dt = MAX_DT" (what /root/reference/src/spectral/spectral_time_integration.c:14-20 does with its
 * `if (vel > max_vel)`); as ordered bits NaN and +Inf compare above every finite value, so the atomicMax of the
 * reduction carries the flag to the CFL kernel */
__device__ __forceinline__ double nsb_max_sticky(double a, double b) { return (b > a || b != b) ? b : a; }

__device__ __forceinline__ int nsb_band(int j, int ny) { return (j <= ny / 2) ? j : j - ny; }
__device__ __forceinline__ int nsb_iabs(int a) { return a < 0 ? -a : a; }

__device__ __forceinline__ double nsb_ipow(double r, int e) {
    double acc = 1.0, b = r;
    while (e > 0) {
        if (e & 1) acc *= b;
        b *= b;
        e >>= 1;
    }
    return acc;
}

/* hyperviscosity.c:89-98 : w = c*((|k|-kc)*inv_bw)^order above the cutoff */
__device__ __forceinline__ double nsb_hv_weight(const NsbHv &h, double kmag) {
    if (kmag > h.kc) return h.coef * nsb_ipow((kmag - h.kc) * h.inv_bw, h.order);
    return 0.0;
}

/* hyperviscosity.c:124-149 : damping = w*state + sfd_strength*ratio^order*(state - sfd) */
__device__ __forceinline__ cd nsb_damping(const NsbHv &h, double k2, cd W, const cd *SFD, size_t idx) {
    cd damp = cmk(0.0, 0.0);
    const double kmag = __dsqrt_rn(k2);
    if (h.visc) {
        const double w = nsb_hv_weight(h, kmag);
        damp.x = __dmul_rn(w, W.x);
        damp.y = __dmul_rn(w, W.y);
    }
    if (h.sfd) {
        double ratio = 0.0;
        if (h.inv_bw > 0.0 && kmag > h.kc) {
            double sc = (kmag - h.kc) * h.inv_bw;
            sc = sc < 0.0 ? 0.0 : (sc > 1.0 ? 1.0 : sc);
            ratio = nsb_ipow(sc, h.order);
        }
        const double sw = h.sfd_strength * ratio;
        if (sw > 0.0) {
            const cd S = SFD[idx];
            damp.x = __dadd_rn(damp.x, __dmul_rn(sw, __dsub_rn(W.x, S.x)));
            damp.y = __dadd_rn(damp.y, __dmul_rn(sw, __dsub_rn(W.y, S.y)));
        }
    }
    return damp;
}

/*
 * RHS + stage update for one mode. spectral_time_integration.c:50-53 (RHS),
 * hyperviscosity.c:104-150 (damping), SURVEY 8c (+f), :57-60/:70-73/:83-86 (stage
 * states), :96-100 (final combine; ACC carries k1+2k2+2k3 in the reference's
 * summation order), hyperviscosity.c:152-178 (SFD filter update).
 * Explicit _rn intrinsics keep nvcc from contracting into FMAs so the pointwise
 * arithmetic rounds like the -ffp-contract=off oracle.
 */
/* pure arithmetic of one RK mode update: inputs already in registers.
 * Returns the next state (Wst for stages 1-3, the new omega_hat for stage 4). */
__device__ __forceinline__ cd nsb_rk_math(const NsbRk &r, double dt, cd W0v, cd Wstv, cd ACCv, cd Sv, double kxv,
                                          double kyv, int i, int band, cd Nh, cd *acc_out, cd *sfd_out) {
    const cd W = (r.stage == 1) ? W0v : Wstv;
    const double k2 = __dadd_rn(__dmul_rn(kxv, kxv), __dmul_rn(kyv, kyv));
    cd k;
    k.x = __dsub_rn(Nh.x, __dmul_rn(__dmul_rn(k2, W.x), r.inv_Re));
    k.y = __dsub_rn(Nh.y, __dmul_rn(__dmul_rn(k2, W.y), r.inv_Re));
    if (r.hv.visc || r.hv.sfd) {
        const cd damp = nsb_damping(r.hv, k2, W, &Sv, 0);
        k.x = __dsub_rn(k.x, damp.x);
        k.y = __dsub_rn(k.y, damp.y);
    }
    if (r.FB != nullptr && i <= r.ib && nsb_iabs(band) <= r.jb) {
        const cd f = r.FB[(size_t)(band + r.jb) * (r.ib + 1) + i];
        k.x = __dadd_rn(k.x, f.x);
        k.y = __dadd_rn(k.y, f.y);
    }
    *sfd_out = Sv;
    if (r.stage == 1) {
        *acc_out = k;
        const double c = __dmul_rn(0.5, dt);
        return cmk(__dadd_rn(W0v.x, __dmul_rn(c, k.x)), __dadd_rn(W0v.y, __dmul_rn(c, k.y)));
    } else if (r.stage == 2 || r.stage == 3) {
        *acc_out = cmk(__dadd_rn(ACCv.x, __dmul_rn(2.0, k.x)), __dadd_rn(ACCv.y, __dmul_rn(2.0, k.y)));
        const double c = (r.stage == 2) ? __dmul_rn(0.5, dt) : dt;
        return cmk(__dadd_rn(W0v.x, __dmul_rn(c, k.x)), __dadd_rn(W0v.y, __dmul_rn(c, k.y)));
    }
    *acc_out = ACCv;
    const double c = __ddiv_rn(dt, 6.0);
    const cd Wn = cmk(__dadd_rn(W0v.x, __dmul_rn(c, __dadd_rn(ACCv.x, k.x))),
                      __dadd_rn(W0v.y, __dmul_rn(c, __dadd_rn(ACCv.y, k.y))));
    if (r.hv.sfd && r.hv.sfd_strength > 0.0 && r.hv.sfd_timescale > 0.0 && dt > 0.0) {
        double al = __ddiv_rn(dt, r.hv.sfd_timescale);
        al = al > 1.0 ? 1.0 : (al < 0.0 ? 0.0 : al);
        const double om = __dsub_rn(1.0, al);
        *sfd_out = cmk(__dadd_rn(__dmul_rn(om, Sv.x), __dmul_rn(al, Wn.x)),
                       __dadd_rn(__dmul_rn(om, Sv.y), __dmul_rn(al, Wn.y)));
    }
    return Wn;
}

__device__ __forceinline__ cd nsb_rk_update(const NsbRk &r, size_t idx, double kxv, double kyv, int i,
                                            int band, cd Nh) {
    const cd W0v = r.W0[idx];
    const cd Wstv = (r.stage == 1) ? W0v : r.Wst[idx];
    const cd ACCv = (r.stage == 1) ? cmk(0.0, 0.0) : r.ACC[idx];
    const cd Sv = (r.hv.sfd) ? r.SFD[idx] : cmk(0.0, 0.0);
    cd acc, sfd;
    const cd nxt = nsb_rk_math(r, *r.dt, W0v, Wstv, ACCv, Sv, kxv, kyv, i, band, Nh, &acc, &sfd);
    if (r.stage < 4) {
        r.ACC[idx] = acc;
        r.Wst[idx] = nxt;
    } else {
        r.W0[idx] = nxt;
        if (r.hv.sfd) r.SFD[idx] = sfd;
    }
    return nxt;
}

/* ------------------------------------------------------------------ alpha_inv */

enum { NSB_SRC_ARRAYS = 0, NSB_SRC_STEP = 1, NSB_SRC_NL = 2, NSB_SRC_VEL = 3 };

struct AlphaInvArgs {
    int nf;
    const cd *src[4];
    cd *out;
    size_t out_fs;
    int out_pitch;
};

/* spectral_core.c:17-18 : psi = -w/k2, 0 where k2 <= 1e-12 */
__device__ __forceinline__ cd nsb_psi_from_w(cd W, double k2) {
    if (k2 > 1e-12) return cmk(__ddiv_rn(-W.x, k2), __ddiv_rn(-W.y, k2));
    return cmk(0.0, 0.0);
}
/* (i k) * z  (spectral_core.c:30-31,66,76,86,96) */
__device__ __forceinline__ cd nsb_ik(double k, cd z) { return cmk(-(k * z.y), k * z.x); }

template <int MODE>
__device__ __forceinline__ cd nsb_alpha_source(const NsbGeom &g, const AlphaInvArgs &a, int f, int j, int i) {
    const size_t idx = (size_t)j * g.nkx + i;
    if (MODE == NSB_SRC_ARRAYS) return a.src[f][idx];
    const double kxv = __ldg(&g.kx[i]), kyv = __ldg(&g.ky[j]);
    if (MODE == NSB_SRC_NL) { /* stored psi_hat (src[0]) and omega_hat (src[1]) */
        const cd z = (f < 2) ? a.src[0][idx] : a.src[1][idx];
        return nsb_ik((f & 1) ? kyv : kxv, z);
    }
    const cd W = a.src[0][idx];
    const double k2 = __dadd_rn(__dmul_rn(kxv, kxv), __dmul_rn(kyv, kyv));
    const cd psi = nsb_psi_from_w(W, k2);
    if (MODE == NSB_SRC_STEP) {
        const cd z = (f < 2) ? psi : W;
        return nsb_ik((f & 1) ? kyv : kxv, z);
    }
    /* NSB_SRC_VEL: u_hat = i ky psi, v_hat = -i kx psi, omega_hat */
    if (f == 0) return nsb_ik(kyv, psi);
    if (f == 1) { const cd t = nsb_ik(kxv, psi); return cmk(-t.x, -t.y); }
    return W;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_alpha_inv(NsbGeom g, AlphaInvArgs a) {
    NSB_DYN_SMEM(cd, s);
    const int j2 = blockIdx.y;
    const int ic0 = blockIdx.x * g.C;
    const int items = g.Y1 * g.C;
    const int wmul = g.ny / g.Y1;
    for (int f = 0; f < a.nf; ++f) {
        for (int w = threadIdx.x; w < items; w += blockDim.x) {
            const int j1 = w / g.C, c = w - j1 * g.C;
            const int j = j1 * g.Y2 + j2, ic = ic0 + c;
            cd v = cmk(0.0, 0.0);
            if (ic < g.ncol && nsb_iabs(nsb_band(j, g.ny)) <= g.ycut) v = nsb_alpha_source<MODE>(g, a, f, j, ic);
            s[w] = v;
        }
        __syncthreads();
        nsb_fft_smem_dif<true>(s, g.Y1, g.C, g.C, 1, g.WY, wmul);
        cd *out = a.out + (size_t)f * a.out_fs;
        for (int w = threadIdx.x; w < items; w += blockDim.x) {
            const int p = w / g.C, c = w - p * g.C;
            const int ic = ic0 + c;
            if (ic >= g.ncol) continue;
            const int y1 = nsb_digitrev(p, g.Y1);
            cd v = s[w];
            if (g.Y2 > 1) v = cmulc(v, __ldg(&g.WY[(size_t)j2 * y1]));
            out[(size_t)(y1 * g.Y2 + j2) * a.out_pitch + ic] = v;
        }
        __syncthreads();
    }
}

/* ------------------------------------------------------------------ alpha_fwd */

enum { NSB_EPI_STORE = 0, NSB_EPI_NEG = 1, NSB_EPI_RK = 2 };

struct AlphaFwdArgs {
    const cd *in;
    int in_pitch;
    cd *out; /* STORE / NEG: reference-layout spectral array */
    NsbRk rk;
};

template <int EPI>
__global__ void __launch_bounds__(256) k_alpha_fwd(NsbGeom g, AlphaFwdArgs a) {
    NSB_DYN_SMEM(cd, s);
    const int j2 = blockIdx.y;
    const int ic0 = blockIdx.x * g.C;
    const int items = g.Y1 * g.C;
    const int wmul = g.ny / g.Y1;
    for (int w = threadIdx.x; w < items; w += blockDim.x) {
        const int y1 = w / g.C, c = w - y1 * g.C;
        const int ic = ic0 + c;
        cd v = cmk(0.0, 0.0);
        if (ic < g.ncol) {
            v = a.in[(size_t)(y1 * g.Y2 + j2) * a.in_pitch + ic];
            if (g.Y2 > 1) v = cmul(v, __ldg(&g.WY[(size_t)j2 * y1]));
        }
        s[w] = v;
    }
    __syncthreads();
    nsb_fft_smem_dif<false>(s, g.Y1, g.C, g.C, 1, g.WY, wmul);
    for (int w = threadIdx.x; w < items; w += blockDim.x) {
        const int p = w / g.C, c = w - p * g.C;
        const int ic = ic0 + c;
        if (ic >= g.ncol) continue;
        const int j = nsb_digitrev(p, g.Y1) * g.Y2 + j2;
        const int band = nsb_band(j, g.ny);
        const bool keep = nsb_iabs(band) <= g.ycut;
        const size_t idx = (size_t)j * g.nkx + ic;
        const cd v = s[w];
        if (EPI == NSB_EPI_STORE) {
            /* spectral_transforms.c:13-21 : times 1/(nx ny), then mask */
            a.out[idx] = keep ? cmk(v.x * g.norm, v.y * g.norm) : cmk(0.0, 0.0);
        } else if (EPI == NSB_EPI_NEG) {
            /* spectral_core.c:106-111 : nonlinear_hat = -(mask * FFT[J] / (nx ny)) */
            a.out[idx] = keep ? cmk(-(v.x * g.norm), -(v.y * g.norm)) : cmk(0.0, 0.0);
        } else {
            if (!keep) continue; /* masked modes: k_rk_masked, only when they are non-zero */
            const cd Nh = cmk(-(v.x * g.norm), -(v.y * g.norm));
            nsb_rk_update(a.rk, idx, __ldg(&g.kx[ic]), __ldg(&g.ky[j]), ic, band, Nh);
        }
    }
}

/* RK update of the modes the 2/3 mask removes (N_hat == 0 there). Linear, so it is
 * skipped entirely while those modes are all zero (ctx->masked_nonzero == 0). */
__global__ void __launch_bounds__(256) k_rk_masked(NsbGeom g, NsbRk rk) {
    const size_t M = (size_t)g.nkx * g.ny;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < M;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(idx / g.nkx), i = (int)(idx - (size_t)j * g.nkx);
        const int band = nsb_band(j, g.ny);
        if (i <= g.imax && nsb_iabs(band) <= g.ycut) continue;
        nsb_rk_update(rk, idx, __ldg(&g.kx[i]), __ldg(&g.ky[j]), i, band, cmk(0.0, 0.0));
    }
}

/* ----------------------------------------------------------------------- beta */

struct BetaArgs {
    const cd *in;
    cd *out;
    int nf;
};

/* number of direct (k1) and mirror (k1') column segments that intersect the retained set */
NSB_HD inline void nsb_beta_segs(int imax, int X1, int X2, int *nd, int *nm) {
    const int half = (X1 == 1) ? 1 : X1 / 2;
    int d = imax / X2 + 1;
    if (d > half) d = half;
    int m = 0;
    if (X1 > 1 && imax >= X2 / 2) {
        m = (imax - X2 / 2) / X2 + 1;
        if (m > X1 / 2) m = X1 / 2;
    }
    *nd = d;
    *nm = m;
}

/* column held by segment `seg`, lane cc of batch kb; -1 if outside the retained set */
__device__ __forceinline__ int nsb_beta_col(const NsbGeom &g, int seg, int nd, int kb, int cc, bool fwd_write) {
    const int k2 = kb + cc;
    if (k2 > g.X2h) return -1;
    int col;
    if (seg < nd) {
        col = seg * g.X2 + k2;
    } else {
        if (fwd_write && (k2 == 0 || k2 == g.X2h)) return -1; /* duplicates of direct columns */
        col = (seg - nd) * g.X2 + g.X2 - k2;
    }
    return (col <= g.imax) ? col : -1;
}

/*
 * beta_inv<Y2,X1>: grid (k2 batches, Y1, nf). For fixed y1:
 *   step 1  length-Y2 inverse DFT over j2 of T[f][y1][j2][col] for every needed column
 *   step 2  per (y2,k2): gather k1 = 0..X1-1 (k < nx/2 direct, else conj of column nx-k),
 *           drop Im of the DC bin (FFTW c2r convention), length-X1 inverse DFT,
 *           times exp(+2 pi i k2 x1/nx)        -> H[f][y1 + Y1*y2][x1][k2]
 */
template <int Y2, int X1>
__global__ void __launch_bounds__(256) k_beta_inv(NsbGeom g, BetaArgs a) {
    NSB_DYN_SMEM(cd, tile);
    const int kb = blockIdx.x * g.BC;
    const int y1 = blockIdx.y;
    const int f = blockIdx.z;
    int nd, nm; /* direct segments k1 = 0..nd-1, mirror segments k1' = 0..nm-1 */
    nsb_beta_segs(g.imax, X1, g.X2, &nd, &nm);
    const int nseg = nd + nm;
    const cd *in = a.in + (size_t)f * g.fsT;
    cd *out = a.out + (size_t)f * g.fsH;
    /* step 1 */
    for (int w = threadIdx.x; w < nseg * g.BC; w += blockDim.x) {
        const int seg = w / g.BC, cc = w - seg * g.BC;
        const int col = nsb_beta_col(g, seg, nd, kb, cc, false);
        cd x[Y2];
        if (col >= 0) {
#pragma unroll
            for (int j2 = 0; j2 < Y2; ++j2) x[j2] = in[(size_t)(y1 * Y2 + j2) * g.PT + col];
            nsb_dft<Y2, true>(x);
        } else {
#pragma unroll
            for (int j2 = 0; j2 < Y2; ++j2) x[j2] = cmk(0.0, 0.0);
        }
#pragma unroll
        for (int y2 = 0; y2 < Y2; ++y2) tile[(size_t)(y2 * nseg + seg) * g.BC + cc] = x[y2];
    }
    __syncthreads();
    /* step 2 */
    for (int w = threadIdx.x; w < Y2 * g.BC; w += blockDim.x) {
        const int y2 = w / g.BC, cc = w - y2 * g.BC;
        const int k2 = kb + cc;
        if (k2 > g.X2h) continue;
        if (X1 == 1 && k2 > g.imax) continue;
        cd x[X1];
#pragma unroll
        for (int k1 = 0; k1 < X1; ++k1) {
            if (X1 == 1 || k1 < X1 / 2) {
                x[k1] = (k1 < nd) ? tile[(size_t)(y2 * nseg + k1) * g.BC + cc] : cmk(0.0, 0.0);
            } else {
                const int kp = X1 - 1 - k1;
                x[k1] = (kp < nm) ? cconj(tile[(size_t)(y2 * nseg + nd + kp) * g.BC + cc]) : cmk(0.0, 0.0);
            }
        }
        if (k2 == 0) x[0].y = 0.0;
        nsb_dft<X1, true>(x);
        const int y = y1 + g.Y1 * y2;
#pragma unroll
        for (int x1 = 0; x1 < X1; ++x1) {
            cd v = x[x1];
            if (X1 > 1) v = cmulc(v, __ldg(&g.WX[(size_t)k2 * x1]));
            out[((size_t)y * X1 + x1) * g.PH + k2] = v;
        }
    }
}

/*
 * beta_fwd<Y2,X1>: grid (k2 batches, Y1). Mirror of beta_inv for one field:
 *   step 1  per (y2,k2): H'[y][x1][k2] * exp(-2 pi i k2 x1/nx), length-X1 forward DFT;
 *           output m is column m*X2+k2, conj(output X1-1-m') is column m'*X2+X2-k2
 *   step 2  length-Y2 forward DFT over y2              -> T'[y1][j2][col]
 */
template <int Y2, int X1>
__global__ void __launch_bounds__(256) k_beta_fwd(NsbGeom g, BetaArgs a) {
    NSB_DYN_SMEM(cd, tile);
    const int kb = blockIdx.x * g.BC;
    const int y1 = blockIdx.y;
    int nd, nm;
    nsb_beta_segs(g.imax, X1, g.X2, &nd, &nm);
    const int nseg = nd + nm;
    for (int w = threadIdx.x; w < Y2 * g.BC; w += blockDim.x) {
        const int y2 = w / g.BC, cc = w - y2 * g.BC;
        const int k2 = kb + cc;
        cd x[X1];
        const bool live = (k2 <= g.X2h) && !(X1 == 1 && k2 > g.imax);
        if (live) {
            const int y = y1 + g.Y1 * y2;
#pragma unroll
            for (int x1 = 0; x1 < X1; ++x1) {
                cd v = a.in[((size_t)y * X1 + x1) * g.PH + k2];
                if (X1 > 1) v = cmul(v, __ldg(&g.WX[(size_t)k2 * x1]));
                x[x1] = v;
            }
            nsb_dft<X1, false>(x);
        } else {
#pragma unroll
            for (int x1 = 0; x1 < X1; ++x1) x[x1] = cmk(0.0, 0.0);
        }
#pragma unroll
        for (int m = 0; m < X1; ++m) {
            if (X1 == 1 || m < X1 / 2) {
                if (m < nd) tile[(size_t)(y2 * nseg + m) * g.BC + cc] = x[m];
            } else {
                const int kp = X1 - 1 - m;
                if (kp < nm) tile[(size_t)(y2 * nseg + nd + kp) * g.BC + cc] = cconj(x[m]);
            }
        }
    }
    __syncthreads();
    for (int w = threadIdx.x; w < nseg * g.BC; w += blockDim.x) {
        const int seg = w / g.BC, cc = w - seg * g.BC;
        const int col = nsb_beta_col(g, seg, nd, kb, cc, true);
        if (col < 0) continue;
        cd x[Y2];
#pragma unroll
        for (int y2 = 0; y2 < Y2; ++y2) x[y2] = tile[(size_t)(y2 * nseg + seg) * g.BC + cc];
        nsb_dft<Y2, false>(x);
#pragma unroll
        for (int j2 = 0; j2 < Y2; ++j2) a.out[(size_t)(y1 * Y2 + j2) * g.PT + col] = x[j2];
    }
}

/* ---------------------------------------------------------------------- gamma */

enum { NSB_GAMMA_STEP = 0, NSB_GAMMA_C2R = 1, NSB_GAMMA_R2C = 2 };

struct GammaArgs {
    const cd *in;          /* H (STEP: 4 fields, C2R: nf fields) */
    int nvalid_in;         /* valid k2 entries per comb row of `in` */
    cd *out;               /* H' (STEP, R2C) */
    int nvalid_out;
    int nf;                /* C2R: number of fields (1..4) */
    double *phys_out[4];   /* C2R: nullable per field */
    const double *phys_in; /* R2C */
    unsigned long long *umax_bits; /* nullable: max over points of (f0^2+f1^2) as ordered bits */
    double *sums;          /* nullable: sums[f] += sum of squares of field f (C2R) */
};

__device__ __forceinline__ void nsb_block_max_to_global(double v, unsigned long long *dst, double *red) {
    /* red: blockDim doubles of smem. All threads call. */
    red[threadIdx.x] = v;
    __syncthreads();
    for (int st = blockDim.x >> 1; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) {
            red[threadIdx.x] = nsb_max_sticky(red[threadIdx.x], red[threadIdx.x + st]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicMax(dst, (unsigned long long)__double_as_longlong(red[0]));
    __syncthreads();
}

__device__ __forceinline__ void nsb_block_sum_to_global(double v, double *dst, double *red) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int st = blockDim.x >> 1; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(dst, red[0]);
    __syncthreads();
}

/*
 * gamma<MODE>: one CTA per pair of combs (q0 = 2*blockIdx.x, q1 = q0+1), comb q =
 * (y, x1) = (q / X1, q % X1) holding physical points x = x1 + X1*x2.
 * smem: 4 sequences of length X2 (pitch SB = X2+1) + blockDim doubles.
 *   sequence 2*c+pp = packed pair pp (fields 2pp, 2pp+1) of comb c.
 */
template <int MODE>
__global__ void __launch_bounds__(256) k_gamma(NsbGeom g, GammaArgs a) {
    NSB_DYN_SMEM(cd, s);
    const int X2 = g.X2, X2h = g.X2h, SB = X2 + 1;
    double *red = reinterpret_cast<double *>(s + 4 * (size_t)SB);
    const int q0 = 2 * blockIdx.x;
    const int wmul = g.nx / X2;
    const int npair = (MODE == NSB_GAMMA_STEP) ? 2 : (a.nf + 1) / 2;
    const int nfields = (MODE == NSB_GAMMA_STEP) ? 4 : a.nf;

    if (MODE != NSB_GAMMA_R2C) {
        /* Hermitian fill: Z[k] = A[k] + i B[k], Z[X2-k] = conj(A[k]) + i conj(B[k]);
         * Im of the DC and Nyquist bins ignored as FFTW's c2r does. */
        const int per = X2h + 1;
        for (int w = threadIdx.x; w < 2 * npair * per; w += blockDim.x) {
            const int b = w / per, k = w - b * per;
            const int comb = b / npair, pp = b - comb * npair;
            const size_t row = (size_t)(q0 + comb) * g.PH;
            cd A = cmk(0.0, 0.0), B = cmk(0.0, 0.0);
            if (k < a.nvalid_in) {
                A = a.in[(size_t)(2 * pp) * g.fsH + row + k];
                if (2 * pp + 1 < nfields) B = a.in[(size_t)(2 * pp + 1) * g.fsH + row + k];
            }
            cd *seq = s + (size_t)(2 * comb + pp) * SB;
            if (k == 0 || k == X2h) {
                seq[k] = cmk(A.x, B.x);
            } else {
                seq[k] = cmk(A.x - B.y, A.y + B.x);
                seq[X2 - k] = cmk(A.x + B.y, B.x - A.y);
            }
        }
        __syncthreads();
        /* sequences used: comb c, pair pp at slot 2c+pp; with npair == 1 slots 0 and 2 */
        if (npair == 2) {
            nsb_fft_smem_dif<true>(s, X2, 4, 1, SB, g.WX, wmul);
        } else {
            nsb_fft_smem_dif<true>(s, X2, 2, 1, 2 * SB, g.WX, wmul);
        }
    }

    if (MODE == NSB_GAMMA_STEP) {
        double umax = 0.0;
        for (int p = threadIdx.x; p < X2; p += blockDim.x) {
            const cd a0 = s[p], b0 = s[SB + p];
            const cd a1 = s[2 * SB + p], b1 = s[3 * SB + p];
            /* spectral_core.c:103 : J = psi_x*omega_y - psi_y*omega_x */
            const double J0 = __dsub_rn(__dmul_rn(a0.x, b0.y), __dmul_rn(a0.y, b0.x));
            const double J1 = __dsub_rn(__dmul_rn(a1.x, b1.y), __dmul_rn(a1.y, b1.x));
            /* spectral_time_integration.c:16 : |u|^2 = psi_y^2 + psi_x^2 */
            const double m0 = __dadd_rn(__dmul_rn(a0.y, a0.y), __dmul_rn(a0.x, a0.x));
            const double m1 = __dadd_rn(__dmul_rn(a1.y, a1.y), __dmul_rn(a1.x, a1.x));
            umax = nsb_max_sticky(umax, nsb_max_sticky(m0, m1));
            s[p] = cmk(J0, J1);
        }
        __syncthreads();
        if (a.umax_bits != nullptr) nsb_block_max_to_global(umax, a.umax_bits, red);
    } else if (MODE == NSB_GAMMA_C2R) {
        double umax = 0.0, sq[4] = {0.0, 0.0, 0.0, 0.0};
        for (int w = threadIdx.x; w < 2 * X2; w += blockDim.x) {
            const int comb = w / X2, p = w - comb * X2;
            const int q = q0 + comb;
            const int y = q / g.X1, x1 = q - y * g.X1;
            const size_t o = (size_t)y * g.nx + x1 + (size_t)g.X1 * nsb_digitrev(p, X2);
            for (int pp = 0; pp < npair; ++pp) {
                const cd z = s[(size_t)(2 * comb + pp) * SB + p];
                if (a.phys_out[2 * pp] != nullptr) a.phys_out[2 * pp][o] = z.x;
                sq[2 * pp] += z.x * z.x;
                if (2 * pp + 1 < nfields) {
                    if (a.phys_out[2 * pp + 1] != nullptr) a.phys_out[2 * pp + 1][o] = z.y;
                    sq[2 * pp + 1] += z.y * z.y;
                }
                if (pp == 0) umax = nsb_max_sticky(umax, __dadd_rn(__dmul_rn(z.x, z.x), __dmul_rn(z.y, z.y)));
            }
        }
        __syncthreads();
        if (a.umax_bits != nullptr) nsb_block_max_to_global(umax, a.umax_bits, red);
        if (a.sums != nullptr)
            for (int f = 0; f < nfields; ++f) nsb_block_sum_to_global(sq[f], a.sums + f, red);
        return;
    } else { /* R2C: gather the two combs into one packed sequence, digit-reversed */
        for (int p = threadIdx.x; p < X2; p += blockDim.x) {
            const int x2 = nsb_digitrev(p, X2);
            double v[2];
            for (int comb = 0; comb < 2; ++comb) {
                const int q = q0 + comb;
                const int y = q / g.X1, x1 = q - y * g.X1;
                v[comb] = a.phys_in[(size_t)y * g.nx + x1 + (size_t)g.X1 * x2];
            }
            s[p] = cmk(v[0], v[1]);
        }
        __syncthreads();
    }

    /* forward transform of the packed pair (comb0 + i comb1), then split */
    nsb_fft_smem_dit<false>(s, X2, 1, 1, SB, g.WX, wmul);
    for (int k = threadIdx.x; k < a.nvalid_out; k += blockDim.x) {
        const cd P = s[k], Q = s[(X2 - k) & (X2 - 1)];
        a.out[(size_t)q0 * g.PH + k] = cmk(0.5 * (P.x + Q.x), 0.5 * (P.y - Q.y));
        a.out[(size_t)(q0 + 1) * g.PH + k] = cmk(0.5 * (P.y + Q.y), -0.5 * (P.x - Q.x));
    }
}

/* ------------------------------------------------------------ pointwise kernels */

/* spectral_core.c:8-34 over all modes: psi_hat, u_hat, v_hat from omega_hat */
__global__ void __launch_bounds__(256) k_velocity_from_vorticity(NsbGeom g, const cd *W, cd *psi, cd *uh, cd *vh) {
    const size_t M = (size_t)g.nkx * g.ny;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < M;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(idx / g.nkx), i = (int)(idx - (size_t)j * g.nkx);
        const double kxv = g.kx[i], kyv = g.ky[j];
        const double k2 = __dadd_rn(__dmul_rn(kxv, kxv), __dmul_rn(kyv, kyv));
        const cd p = nsb_psi_from_w(W[idx], k2);
        psi[idx] = p;
        uh[idx] = nsb_ik(kyv, p);
        const cd t = nsb_ik(kxv, p);
        vh[idx] = cmk(-t.x, -t.y);
    }
}

/* spectral_transforms.c:35-43 */
__global__ void __launch_bounds__(256) k_apply_dealiasing(NsbGeom g, cd *f) {
    const size_t M = (size_t)g.nkx * g.ny;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < M;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(idx / g.nkx), i = (int)(idx - (size_t)j * g.nkx);
        if (!(i <= g.imax && nsb_iabs(nsb_band(j, g.ny)) <= g.ycut)) f[idx] = cmk(0.0, 0.0);
    }
}

/* flag |= 1 if any mode removed by the mask is non-zero */
__global__ void __launch_bounds__(256) k_scan_masked(NsbGeom g, const cd *f, unsigned *flag) {
    const size_t M = (size_t)g.nkx * g.ny;
    bool hit = false;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < M;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(idx / g.nkx), i = (int)(idx - (size_t)j * g.nkx);
        if (i <= g.imax && nsb_iabs(nsb_band(j, g.ny)) <= g.ycut) continue;
        const cd v = f[idx];
        if (v.x != 0.0 || v.y != 0.0) hit = true;
    }
    if (hit) atomicOr(flag, 1u);
}

/* hyperviscosity.c:104-150 as a standalone entry point */
__global__ void __launch_bounds__(256) k_hv_apply(NsbGeom g, NsbHv hv, cd *rhs, const cd *state, const cd *sfd) {
    const size_t M = (size_t)g.nkx * g.ny;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < M;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(idx / g.nkx), i = (int)(idx - (size_t)j * g.nkx);
        const double kxv = g.kx[i], kyv = g.ky[j];
        const double k2 = __dadd_rn(__dmul_rn(kxv, kxv), __dmul_rn(kyv, kyv));
        const cd d = nsb_damping(hv, k2, state[idx], sfd, idx);
        const cd r = rhs[idx];
        rhs[idx] = cmk(__dsub_rn(r.x, d.x), __dsub_rn(r.y, d.y));
    }
}

/* hyperviscosity.c:152-178 as a standalone entry point */
__global__ void __launch_bounds__(256) k_sfd_post(NsbGeom g, double alpha, cd *sfd, const cd *state) {
    const size_t M = (size_t)g.nkx * g.ny;
    const double om = 1.0 - alpha;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < M;
         idx += (size_t)gridDim.x * blockDim.x) {
        const cd S = sfd[idx], W = state[idx];
        sfd[idx] = cmk(__dadd_rn(__dmul_rn(om, S.x), __dmul_rn(alpha, W.x)),
                       __dadd_rn(__dmul_rn(om, S.y), __dmul_rn(alpha, W.y)));
    }
}

/* max over points of u^2+v^2 from physical arrays (EAGER-mode CFL, spectral_time_integration.c:14-20) */
__global__ void __launch_bounds__(256) k_max_speed2(const double *u, const double *v, size_t n,
                                                    unsigned long long *umax_bits) {
    NSB_DYN_SMEM(double, red);
    double m = 0.0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        m = fmax(m, __dadd_rn(__dmul_rn(u[k], u[k]), __dmul_rn(v[k], v[k])));
    nsb_block_max_to_global(m, umax_bits, red);
}

struct NsbScalars {
    double dt;                     /* time step used by the RK update */
    double t;                      /* simulation time */
    unsigned long long umax_bits;  /* max |u|^2 as ordered bits */
    double sums[4];                /* diagnostic sums of squares */
    unsigned masked_nonzero;
    int step;
    double dt_hist[256];
    unsigned peer_flags[8];        /* device barrier of the slab decomposition: epoch announced by each peer */
    double stats[8];               /* nsb_spectral_stats accumulators */
    unsigned nonfinite;            /* set by the CFL kernel when max |u|^2 is NaN or Inf */
    int nonfinite_step;            /* step counter at that moment */
};

/* spectral_time_integration.c:22-30 : dt = clamp(min(cfl*dx/(|u|max+1e-12), 0.5*Re*dx^2)) */
__global__ void k_cfl_dt(NsbScalars *sc, double cfl, double dx_min, double Re, double min_dt, double max_dt) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (sc->umax_bits >= 0x7FF0000000000000ull && !sc->nonfinite) {
        sc->nonfinite = 1u;
        sc->nonfinite_step = sc->step;
    }
    const double max_vel = __dsqrt_rn(__longlong_as_double((long long)sc->umax_bits));
    const double dt_cfl = __ddiv_rn(__dmul_rn(cfl, dx_min), __dadd_rn(max_vel, 1e-12));
    const double dt_visc = __dmul_rn(__dmul_rn(__dmul_rn(0.5, Re), dx_min), dx_min);
    double dt = fmin(dt_cfl, dt_visc);
    dt = fmax(dt, min_dt);
    dt = fmin(dt, max_dt);
    sc->dt = dt;
    sc->umax_bits = 0ull;
}

struct NsbPeerScalars {
    NsbScalars *sc[8];
};

/* the same, with max |u| taken over the slabs of all devices (their scalar blocks are peer memory) */
__global__ void k_cfl_dt_multi(NsbScalars *sc, NsbPeerScalars ps, int world, double cfl, double dx_min, double Re,
                               double min_dt, double max_dt) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long bits = 0ull;
    for (int p = 0; p < world; ++p) {
        const unsigned long long b = *reinterpret_cast<volatile unsigned long long *>(&ps.sc[p]->umax_bits);
        if (b > bits) bits = b;
    }
    if (bits >= 0x7FF0000000000000ull && !sc->nonfinite) {
        sc->nonfinite = 1u;
        sc->nonfinite_step = sc->step;
    }
    const double max_vel = __dsqrt_rn(__longlong_as_double((long long)bits));
    const double dt_cfl = __ddiv_rn(__dmul_rn(cfl, dx_min), __dadd_rn(max_vel, 1e-12));
    const double dt_visc = __dmul_rn(__dmul_rn(__dmul_rn(0.5, Re), dx_min), dx_min);
    double dt = fmin(dt_cfl, dt_visc);
    dt = fmax(dt, min_dt);
    dt = fmin(dt, max_dt);
    sc->dt = dt;
}

/*
 * Device barrier across the processes of a slab-decomposed run: lane p announces `epoch`
 * in peer p's scalar block (NVLink peer store after a system-scope fence, so the T / T'
 * pushes of the preceding kernels are visible first) and then waits for peer p's
 * announcement in the local block.
 */
__global__ void k_peer_barrier(NsbScalars *sc, NsbPeerScalars ps, int rank, int world, unsigned epoch) {
    const int p = threadIdx.x;
    if (p >= world) return;
    nsb_fence_system();
    *reinterpret_cast<volatile unsigned *>(&ps.sc[p]->peer_flags[rank]) = epoch;
    while (*reinterpret_cast<volatile unsigned *>(&sc->peer_flags[p]) < epoch) nsb_backoff();
    nsb_fence_system();
}

__global__ void k_set_dt(NsbScalars *sc, double dt) {
    if (threadIdx.x == 0 && blockIdx.x == 0) sc->dt = dt;
}

/* spectral_time_integration.c:104 : t += dt (plus a small dt history ring for the host) */
__global__ void k_step_end(NsbScalars *sc) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    sc->t = __dadd_rn(sc->t, sc->dt);
    sc->dt_hist[sc->step & 255] = sc->dt;
    sc->step += 1;
}

/* ------------------------------------------------- gamma, X2 == 1024 fast path */

struct GammaFastArgs {
    const cd *in;   /* H, 4 fields */
    cd *out;        /* H' */
    int nvalid_in, nvalid_out;
    unsigned long long *umax_bits;
    int npairs;
};

/* per-warp shared memory: transpose buffer | parked z1 (complex) | parked J of comb 0 |
 * TMA landing buffer (two spectrum rows of 513) | mbarrier */
#define NSB_G1024_T_BYTES (32 * NSB_TP * 8)
#define NSB_G1024_L_BYTES (2 * 513 * 16)
#define NSB_G1024_WARP_BYTES (NSB_G1024_T_BYTES + 1024 * 16 + 1024 * 8 + NSB_G1024_L_BYTES + 16)
#define NSB_G1024_SMEM(warps) (1024 * 16 + (warps) * NSB_G1024_WARP_BYTES)

/* Z[k] = A[k] + i B[k] for k <= 512, Z[k] = conj(A[1024-k]) + i conj(B[1024-k]) above;
 * Im of the DC and Nyquist bins dropped (FFTW c2r convention). Lane t gets k = t + 32 m.
 * A, B: the two spectrum rows in the shared-memory landing buffer (zero beyond nvalid). */
__device__ __forceinline__ void nsb_load_hermitian1024(cd (&x)[32], const cd *A, const cd *B, int nvalid, int lane) {
#pragma unroll
    for (int m = 0; m < 32; ++m) {
        const int k = lane + 32 * m;
        const bool lower = (m < 16) || (m == 16 && lane == 0);
        const int kk = lower ? k : 1024 - k;
        cd av = cmk(0.0, 0.0), bv = cmk(0.0, 0.0);
        if (kk < nvalid) {
            av = A[kk];
            bv = B[kk];
        }
        if (k == 0 || k == 512) x[m] = cmk(av.x, bv.x);
        else if (lower) x[m] = cmk(av.x - bv.y, av.y + bv.x);
        else x[m] = cmk(av.x + bv.y, bv.x - av.y);
    }
}

/*
 * gamma for X2 == 1024: one WARP per comb pair, no block barrier after the twiddle
 * table is staged. Per pair: 4 inverse + 1 forward warp FFT-1024, Jacobian and max|u|^2
 * in registers. The two spectrum rows of the NEXT transform are fetched by TMA bulk copies
 * (cp.async.bulk -> mbarrier) into the warp's landing buffer while the current transform
 * runs, so global latency is off the critical path with one warp per scheduler.
 * Same inputs/outputs as k_gamma<NSB_GAMMA_STEP>.
 */
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_gamma1024(NsbGeom g, GammaFastArgs a) {
    NSB_DYN_SMEM(unsigned char, raw);
    cd *TW = reinterpret_cast<cd *>(raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wb = raw + 1024 * 16 + (size_t)warp * NSB_G1024_WARP_BYTES;
    double *T = reinterpret_cast<double *>(wb);
    cd *P = reinterpret_cast<cd *>(wb + NSB_G1024_T_BYTES);
    double *PJ = reinterpret_cast<double *>(wb + NSB_G1024_T_BYTES + 1024 * 16);
    cd *L = reinterpret_cast<cd *>(wb + NSB_G1024_T_BYTES + 1024 * 16 + 1024 * 8);
    nsb_mbar_t *bar = reinterpret_cast<nsb_mbar_t *>(wb + NSB_G1024_T_BYTES + 1024 * 16 + 1024 * 8 + NSB_G1024_L_BYTES);
    nsb_fill_tw1024(TW, g.TWX);
    if (lane == 0) nsb_mbar_init(bar, 1);
    __syncthreads();
    const unsigned row_bytes = (unsigned)a.nvalid_in * 16u;
    const int stride = gridDim.x * WARPS;
    const int pair0 = blockIdx.x * WARPS + warp;
    /* job = (pair, comb, pp): rows of fields 2pp, 2pp+1 of comb row 2*pair+comb */
    auto issue = [&](int pair, int comb, int pp) {
        if (lane == 0) {
            const size_t row = (size_t)(2 * pair + comb) * g.PH;
            nsb_fence_proxy_async();
            nsb_mbar_expect_tx(bar, 2u * row_bytes);
            nsb_bulk_g2s(L, a.in + (size_t)(2 * pp) * g.fsH + row, row_bytes, bar);
            nsb_bulk_g2s(L + 513, a.in + (size_t)(2 * pp + 1) * g.fsH + row, row_bytes, bar);
        }
    };
    unsigned phase = 0;
    if (pair0 < a.npairs) issue(pair0, 0, 0);
    double umax = 0.0;
    for (int pair = pair0; pair < a.npairs; pair += stride) {
        cd x[32];
#pragma unroll 1
        for (int comb = 0; comb < 2; ++comb) {
            nsb_mbar_wait(bar, phase);
            phase ^= 1u;
            nsb_load_hermitian1024(x, L, L + 513, a.nvalid_in, lane);
            __syncwarp();
            issue(pair, comb, 1);
            nsb_fft1024_warp<true>(x, T, TW, lane);
#pragma unroll
            for (int m = 0; m < 32; ++m) P[m * 32 + lane] = x[m];
            nsb_mbar_wait(bar, phase);
            phase ^= 1u;
            nsb_load_hermitian1024(x, L, L + 513, a.nvalid_in, lane);
            __syncwarp();
            if (comb == 0) issue(pair, 1, 0);
            else if (pair + stride < a.npairs) issue(pair + stride, 0, 0);
            nsb_fft1024_warp<true>(x, T, TW, lane);
#pragma unroll
            for (int m = 0; m < 32; ++m) {
                const cd z1 = P[m * 32 + lane], z2 = x[m];
                /* spectral_core.c:103 and spectral_time_integration.c:16 */
                const double J = __dsub_rn(__dmul_rn(z1.x, z2.y), __dmul_rn(z1.y, z2.x));
                umax = nsb_max_sticky(umax, __dadd_rn(__dmul_rn(z1.y, z1.y), __dmul_rn(z1.x, z1.x)));
                if (comb == 0) PJ[m * 32 + lane] = J;
                else x[m] = cmk(PJ[m * 32 + lane], J);
            }
        }
        nsb_fft1024_warp<false>(x, T, TW, lane);
#pragma unroll
        for (int m = 0; m < 32; ++m) P[m * 32 + lane] = x[m]; /* natural order: index k = lane + 32 m */
        __syncwarp();
        cd *o0 = a.out + (size_t)(2 * pair) * g.PH, *o1 = o0 + g.PH;
#pragma unroll
        for (int m = 0; m <= 16; ++m) {
            const int k = lane + 32 * m;
            if (k <= 512 && k < a.nvalid_out) {
                const cd Pk = P[k], Q = P[(1024 - k) & 1023];
                o0[k] = cmk(0.5 * (Pk.x + Q.x), 0.5 * (Pk.y - Q.y));
                o1[k] = cmk(0.5 * (Pk.y + Q.y), -0.5 * (Pk.x - Q.x));
            }
        }
        __syncwarp();
    }
    if (a.umax_bits != nullptr) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) umax = nsb_max_sticky(umax, __shfl_xor_sync(0xffffffffu, umax, off));
        if (lane == 0) atomicMax(a.umax_bits, (unsigned long long)__double_as_longlong(umax));
    }
}

/* ------------------------------------------- alpha, Y1 == 1024 fused fast path */

/*
 * Internal "K layout" of the RK state arrays used by the fast path: retained columns
 * only, K[(ic*Y2 + j2)*Y1 + j1] with j = j1*Y2 + j2, so that the 32 lanes of a warp that
 * owns column ic touch 512 contiguous bytes. Converted from / to the reference layout
 * (f_hat[j*nkx + i]) only at the API boundary (upload, download, diagnostics).
 */
/* columns [c0, c1) are the ones this device owns (slab decomposition); K is indexed by ic - c0 */
__global__ void __launch_bounds__(256) k_ref_to_klayout(NsbGeom g, const cd *ref, cd *K, int c0, int c1) {
    __shared__ cd tile[32][33];
    /* grid: (ceil((c1-c0)/32), Y1/32, Y2); block (32, 8) */
    const int j2 = blockIdx.z, ic0 = c0 + blockIdx.x * 32, j10 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int j = (j10 + r) * g.Y2 + j2, ic = ic0 + threadIdx.x;
        tile[r][threadIdx.x] = (ic < c1) ? ref[(size_t)j * g.nkx + ic] : cmk(0.0, 0.0);
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int ic = ic0 + r;
        if (ic < c1) K[((size_t)(ic - c0) * g.Y2 + j2) * g.Y1 + j10 + threadIdx.x] = tile[threadIdx.x][r];
    }
}

__global__ void __launch_bounds__(256) k_klayout_to_ref(NsbGeom g, const cd *K, cd *ref, int c0, int c1) {
    __shared__ cd tile[32][33];
    const int j2 = blockIdx.z, ic0 = c0 + blockIdx.x * 32, j10 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int ic = ic0 + r;
        tile[r][threadIdx.x] =
            (ic < c1) ? K[((size_t)(ic - c0) * g.Y2 + j2) * g.Y1 + j10 + threadIdx.x] : cmk(0.0, 0.0);
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int j = (j10 + r) * g.Y2 + j2, ic = ic0 + threadIdx.x;
        if (ic < c1) ref[(size_t)j * g.nkx + ic] = tile[threadIdx.x][r];
    }
}

/*
 * Slab decomposition: this device owns retained columns [c0, c1) in spectral
 * space (K-layout state and T' are indexed by ic - c0) and the y rows
 * [rank*2^row_shift, (rank+1)*2^row_shift) in physical space. The inverse output tile is
 * PUSHED row by row into the T buffer of the device that owns the row (outp[owner], peer
 * memory over NVLink; outp[0] is the local buffer when there is one device), i.e. the
 * global transpose of the 2-D FFT is fused into the store phase of the column kernel.
 */
struct AlphaFastArgs {
    const cd *in;    /* T' (forward input), [ny][in_pitch], local columns */
    int in_pitch;
    cd *outp[NSB_MAXP]; /* T of every device: 4 fields x own rows x out_pitch */
    size_t out_fs;   /* field stride of T (own rows * out_pitch) */
    int out_pitch;
    int row_shift;   /* log2(rows per device) */
    int tma_rl, tma_br; /* TMA stores: y1 rows per device, rows per box */
    int tma_in;      /* the forward tile comes from the T' exchange buffer the tensor map tm.tp describes */
    int c0, c1;      /* own columns */
    const cd *Wsrc;  /* K-layout state feeding the inverse part when there is no forward part */
    NsbRk rk;        /* K-layout pointers */
};

/* tensor maps of the T buffer of every device (3-D: {2*PT doubles, Y2, 4 fields x own y1 rows},
 * 128B swizzle) for the TMA tile stores of k_alpha1024<.., TMA = true> */
struct NsbTmaps {
    nsb_tmap m[NSB_MAXP];
    nsb_tmap tp; /* the local T' buffer: {2*cb doubles, Y2, Y1}, for the TMA load of the forward tile */
    nsb_tmap m4[NSB_MAXP], tp4; /* the same with 4-column boxes and the 64-byte swizzle (two-group column kernel) */
};

/* modes per lane whose RK state loads are in flight together in the in-register RK update of
 * k_alpha1024 (the transform registers stay live, so there is room for about 4 x 3 complex values) */
#ifndef NSB_RK_BATCH
#define NSB_RK_BATCH 4
#endif
/* 1: fully unrolled update on the live transform registers; 2 (default): rolled, double-buffered loop
 * with the values parked in the lane's own tile slots */
#ifndef NSB_RK_MODE
#define NSB_RK_MODE 2
#endif
#ifndef NSB_TILE_UNROLL
#define NSB_TILE_UNROLL 8
#endif
#define NSB_PRAGMA_(x) _Pragma(#x)
#define NSB_UNROLL(n) NSB_PRAGMA_(unroll n)
#define NSB_A1024_COLS 4 /* granularity of the column partition; kernels take 4 or 8 columns per CTA */
#define NSB_A1024_SMEM(COLS) (1024 * 16 + 1024 * (COLS) * 16 + 32 * 16 + 1024 * 8)
#define NSB_A1024_SMEM_TMA (NSB_A1024_SMEM(8) + 8 * 8192 + 32 + 1024)
/* two groups: 2 x 64 KB tiles | 16 KB FFT twiddles | 2 x (TB 512 B + KY 8 KB) | 8 x 8 KB transpose buffers | 2 mbarriers */
#define NSB_A1024_SMEM_2G (2 * 1024 * 4 * 16 + 1024 * 16 + 2 * (32 * 16 + 1024 * 8) + 8 * 8192 + 32 + 1024)
/* 1 (default): the TMA variants (8 warps, one CTA per SM) compute the reciprocal 1/k^2 once per mode: the first
 * inverse job also forms the input of the second and parks it in the lane's own tensor-memory columns */
#ifndef NSB_ALPHA_PARK
#define NSB_ALPHA_PARK 1
#endif
#ifndef NSB_ALPHA_UNPARK16
#define NSB_ALPHA_UNPARK16 0
#endif

/* shared tile S[1024 rows][4 columns], 16-byte chunks XOR-swizzled so that row-wise
 * (cooperative, coalesced) and column-wise (per-warp FFT) accesses are both conflict free */
template <int COLS>
__device__ __forceinline__ int nsb_sidx(int row, int c) {
    return COLS == 4 ? row * 4 + (c ^ ((row >> 1) & 3)) : row * COLS + (c ^ (row & (COLS - 1)));
}

/* warp FFT-1024 whose transpose goes through the warp's own column of S (skewed so the
 * strided read is conflict free as well) */
template <bool INV, int COLS>
__device__ __forceinline__ void nsb_fft1024_warp_col(cd (&x)[32], cd *S, int col, const cd *TW, int lane) {
    nsb_dft32<INV>(x);
#pragma unroll
    for (int r = 1; r < 32; ++r) {
        const cd tw = TW[r * 32 + lane];
        x[r] = INV ? cmulc(x[r], tw) : cmul(x[r], tw);
    }
#pragma unroll
    for (int r = 0; r < 32; ++r) S[nsb_sidx<COLS>(32 * r + ((lane + r) & 31), col)] = x[r];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 32; ++q) x[q] = S[nsb_sidx<COLS>(32 * lane + ((q + lane) & 31), col)];
    __syncwarp();
    nsb_dft32<INV>(x);
}

#ifdef NSB_EMULATE
#define NSB_NOINLINE
#else
#define NSB_NOINLINE __noinline__
#endif

__device__ NSB_NOINLINE cd nsb_rk_update_call(const NsbRk *r, size_t idx, double kxv, double kyv, int i, int band,
                                             cd Nh) {
    return nsb_rk_update(*r, idx, kxv, kyv, i, band, Nh);
}

/*
 * Fused column kernel for Y1 == 1024. CTA = 4 warps = 4 adjacent retained columns at one
 * residue j2; during a transform a warp owns one column and holds its 1024 elements in
 * registers. One transform body (the inverse; the forward one runs as conj o inverse o
 * conj) is looped over up to five jobs:
 *   job 0 (FWD)  T' tile -> cooperative coalesced load into swizzled smem -> y twiddle ->
 *                forward warp FFT -> back to the tile; then ALL threads sweep the tile for
 *                -1/(nx ny), mask, RHS + damping + forcing + RK stage update against the
 *                K-layout state (512-byte coalesced, several loads in flight per thread)
 *   jobs 1-4 (INV) next-stage state (K layout) -> psi, i k fields -> inverse warp FFT ->
 *                y twiddle -> tile -> cooperative coalesced store to T[f]
 * Replaces k_alpha_fwd<RK> of stage s followed by k_alpha_inv<STEP> of stage s+1.
 */
template <bool FWD, bool INV, int COLS, bool TMA, int GROUPS = 1>
__global__ void __launch_bounds__(COLS * 32 * GROUPS, 8 / (COLS * GROUPS))
    k_alpha1024(NsbGeom g, AlphaFastArgs a, const NSB_GRID_CONSTANT NsbTmaps tm) {
    /* GROUPS == 2 (TMA, COLS == 4): the CTA is two independent 4-warp groups, each with its own 64 KB tile, ky /
     * twiddle rows, mbarrier and tile sequence, synchronising on its own named barrier -- two tiles in flight per SM,
     * so that one group's T' load, RK update and store drain overlap the other's transforms. Only the 16 KB
     * FFT twiddle table is shared. (gt = thread index within the group.) */
    const int gi = GROUPS > 1 ? (int)(threadIdx.x / (COLS * 32)) : 0;
    const int gt = GROUPS > 1 ? (int)(threadIdx.x % (COLS * 32)) : (int)threadIdx.x;
#define NSB_GSYNC()                                         \
    do {                                                    \
        if (GROUPS > 1) nsb_group_barrier(1 + gi, COLS * 32); \
        else __syncthreads();                               \
    } while (0)
    NSB_DYN_SMEM(cd, smem_raw);
    /* TMA: the tile must start on a 1024-byte boundary so that nsb_sidx<8> IS the 128B swizzle */
    cd *smem = TMA ? smem_raw + nsb_smem_align_pad(smem_raw, 1024) / sizeof(cd) : smem_raw;
    cd *S = smem + (size_t)gi * 1024 * COLS;
    cd *TW = smem + (size_t)GROUPS * 1024 * COLS;
    cd *TB = TW + 1024 + (size_t)gi * 32; /* TB[m] = exp(-2 pi i (32 j2 m)/ny) */
    const int lane = threadIdx.x & 31, warp = gt >> 5; /* warp = column of the group's tile */
    const int cc = gt & (COLS - 1), rr = gt / COLS; /* cooperative copies: COLS threads per row */
    double *KY = reinterpret_cast<double *>(TW + 1024 + GROUPS * 32) + (size_t)gi * 1024; /* ky of the tile's rows, by j1 */
    /* TMA: private 8 KB transpose buffer per warp, so the tile is not touched while its store drains */
    double *TS = reinterpret_cast<double *>(TW + 1024 + GROUPS * 32) + (size_t)GROUPS * 1024 + (size_t)(threadIdx.x >> 5) * 1024;
    nsb_mbar_t *tbar = reinterpret_cast<nsb_mbar_t *>(reinterpret_cast<double *>(TW + 1024 + GROUPS * 32) +
                                                      (size_t)GROUPS * 1024 + 8 * 1024) + gi; /* TMA only */
    const cd *Wre = FWD ? a.rk.Wst : a.Wsrc; /* state the inverse fields derive from */
    constexpr bool PARK = TMA && INV && (NSB_ALPHA_PARK != 0);
    constexpr bool PARKT = TMA && (NSB_ALPHA_PARK != 0); /* y twiddles ta * TB[m] of the lane in columns 128..255 */
    nsb_fill_tw1024(TW, g.TWY);
    if (TMA && gt == 0) nsb_mbar_init(tbar, 1);
    unsigned tbase = 0u, tcol = 0u;
    if (PARKT) {
        /* 256 tensor-memory columns per warp (two warps share a lane quadrant); block barrier inside */
        tbase = nsb_tmem_alloc_all(reinterpret_cast<unsigned *>(tbar - gi + 2));
        tcol = nsb_tmem_warp_base(tbase, (int)(threadIdx.x >> 5), (int)(threadIdx.x >> 7) * 256);
    } else if (GROUPS > 1) {
        __syncthreads(); /* the shared twiddle table; from here on the groups only meet themselves */
    }
    /*
     * PERSISTENT over tiles (round 2): tile t = (j2, column tile), j2 slowest, CTA b takes t = b, b + gridDim.x, ...
     * The 16 KB twiddle table is staged once per CTA, the ky / y-twiddle rows only when j2 changes, and the last
     * TMA store of a tile drains behind the start of the next tile instead of holding the CTA at EXIT (round-2
     * profile of the one-tile-per-CTA version: 4-11 % of the stall samples at EXIT, 3 % in the table fills).
     * The LDG/STG variants are launched with one tile per CTA (the loop runs once).
     */
    const int ntx = (a.c1 - a.c0 + COLS - 1) / COLS;
    const int ntiles = ntx * g.Y2;
    unsigned tphase = 0u;
    int j2_cur = -1;
    unsigned keepmask = 0u;
    cd ta = cmk(1.0, 0.0);
    cd x[32];
    bool prefetched = false; /* x already holds the input of the next inverse job */
#pragma unroll 1
    for (int t = GROUPS * (int)blockIdx.x + gi; t < ntiles; t += GROUPS * (int)gridDim.x) {
    const int j2 = t / ntx, ic0 = a.c0 + (t - j2 * ntx) * COLS;
    const int ic = ic0 + warp;
    const bool active = ic < a.c1;
    /* the previous tile's last tile store must have READ the tile, and every warp must be done with the tile and
     * the ky / twiddle rows, before they are refilled */
    if (TMA && gt == 0) nsb_tma_store_wait_read();
    NSB_GSYNC();
    if (j2 != j2_cur) {
        j2_cur = j2;
        if (gt < 32) TB[gt] = __ldg(&g.WY[(size_t)32 * j2 * gt]);
        /* y twiddle exp(-2 pi i j2 y1/ny), y1 = lane + 32 m, factored as ta * TB[m] (measured: a full
         * 1024-entry table in shared memory costs more in lost L1 than the extra complex multiply) */
        ta = __ldg(&g.WY[(size_t)j2 * lane]);
        if (PARKT && g.Y2 > 1) {
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                double tpk[8];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const cd w = cmul(ta, __ldg(&g.WY[(size_t)32 * j2 * (4 * h + u)]));
                    tpk[2 * u] = w.x;
                    tpk[2 * u + 1] = w.y;
                }
                nsb_tmem_st8(tcol + 128 + 16 * h, tpk);
            }
            nsb_tmem_wait_st();
        }
        for (int k = gt; k < 1024; k += COLS * 32) KY[k] = __ldg(&g.kyK[(size_t)j2 * 1024 + k]);
        /* rows of this lane that survive the 2/3 mask, one bit per register index (same for every job) */
        keepmask = 0u;
#pragma unroll
        for (int m = 0; m < 32; ++m)
            if (nsb_iabs(nsb_band((lane + 32 * m) * g.Y2 + j2, g.ny)) <= g.ycut) keepmask |= 1u << m;
    }
    const double kxv = active ? __ldg(&g.kx[ic]) : 0.0;
    const size_t kbase = ((size_t)(ic - a.c0) * g.Y2 + j2) * 1024;
    if (FWD && TMA && a.tma_in) {
        /* the whole T' tile (1024 rows x 8 columns) in four tensor loads, in flight while the
         * ky / twiddle rows above are filled by the other threads */
        if (gt == 0) {
            nsb_fence_proxy_async();
            nsb_mbar_expect_tx(tbar, 1024u * COLS * 16u);
            for (int r0 = 0; r0 < 1024; r0 += 256)
                nsb_tma_3d_g2s(S + (size_t)r0 * COLS, COLS == 4 ? &tm.tp4 : &tm.tp, 2 * (ic0 - a.c0), j2, r0, tbar);
        }
    } else if (FWD) {
        const bool colok = ic0 + cc < a.c1;
        NSB_UNROLL(NSB_TILE_UNROLL)
        for (int i = 0; i < 32; ++i) {
            const int row = rr + 32 * i;
            cd v = cmk(0.0, 0.0);
            if (colok) v = a.in[(size_t)(row * g.Y2 + j2) * a.in_pitch + (ic0 - a.c0) + cc];
            S[nsb_sidx<COLS>(row, cc)] = v;
        }
    }
    NSB_GSYNC();
    if (FWD && TMA && a.tma_in) {
        nsb_mbar_wait(tbar, tphase);
        tphase ^= 1u;
    }
    const double kx2 = __dmul_rn(kxv, kxv);
#pragma unroll 1
    for (int job = FWD ? 0 : 1; job <= (INV ? 4 : 0); ++job) {
        if (active) {
            if (job == 0) {
                if (PARKT && g.Y2 > 1) {
                    NsbTmemRaw8 q[2];
                    nsb_tmem_ld8_async(q[0], tcol + 128);
#pragma unroll
                    for (int h = 0; h < 8; ++h) {
                        double tpk[8];
                        nsb_tmem_wait_ld8(tpk, q[h & 1]);
                        if (h < 7) nsb_tmem_ld8_async(q[(h + 1) & 1], tcol + 128 + 16 * (h + 1));
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int m = 4 * h + u;
                            const cd v = cmul(S[nsb_sidx<COLS>(lane + 32 * m, warp)], cmk(tpk[2 * u], tpk[2 * u + 1]));
                            x[m] = cmk(v.x, -v.y);
                        }
                    }
                } else {
#pragma unroll
                for (int m = 0; m < 32; ++m) {
                    cd v = S[nsb_sidx<COLS>(lane + 32 * m, warp)];
                    if (g.Y2 > 1) v = cmul(v, cmul(ta, TB[m]));
                    x[m] = cmk(v.x, -v.y); /* forward = conj o inverse o conj */
                }
                }
                __syncwarp(); /* every lane holds its values before the transpose reuses the slots */
            } else {
                /* PARK: field order 2, 3, 0, 1 -- the state W is parked in the lane's tensor-memory columns by the
                 * first job and read back by the next two (no reload from global memory between jobs); the third
                 * forms -i (kx/k^2) w and -i (ky/k^2) w from ONE reciprocal per mode and parks the latter, over W,
                 * for the fourth. Same formulas per field as the plain order (the compiler contracts the parked twiddle
                 * product into FMAs differently, so T agrees with the unparked variants to rounding, not bit for bit;
                 * every device count runs this same code and stays bit-identical to the others). */
                const int f = PARK ? ((job + 1) & 3) : job - 1;
                /* all 32 state loads in flight at once (unconditional: K layout covers every row);
                 * without PARK they were issued before the previous job's store phase for jobs 2-4 */
                if (!prefetched && !(PARK && job > 1)) {
#pragma unroll
                    for (int m = 0; m < 32; ++m) x[m] = Wre[kbase + lane + 32 * m];
                }
                if (PARK) {
                    /* the parked 32 values back into x, four at a time, the next load in flight behind `use` */
                    auto unpark = [&](auto use) {
#if NSB_ALPHA_UNPARK16 == 2
#pragma unroll
                        for (int h = 0; h < 8; ++h) {
                            double tpk[8];
                            nsb_tmem_ld8(tpk, tcol + 16 * h);
#pragma unroll
                            for (int u = 0; u < 4; ++u) x[4 * h + u] = cmk(tpk[2 * u], tpk[2 * u + 1]);
                            use(h);
                        }
#elif NSB_ALPHA_UNPARK16
#pragma unroll
                        for (int h2 = 0; h2 < 4; ++h2) {
                            double tpk[16];
                            nsb_tmem_ld16(tpk, tcol + 32 * h2);
#pragma unroll
                            for (int u = 0; u < 8; ++u) x[8 * h2 + u] = cmk(tpk[2 * u], tpk[2 * u + 1]);
                            use(2 * h2);
                            use(2 * h2 + 1);
                        }
#else
                        NsbTmemRaw8 q[2];
                        nsb_tmem_ld8_async(q[0], tcol);
#pragma unroll
                        for (int h = 0; h < 8; ++h) {
                            double tpk[8];
                            nsb_tmem_wait_ld8(tpk, q[h & 1]);
                            if (h < 7) nsb_tmem_ld8_async(q[(h + 1) & 1], tcol + 16 * (h + 1));
#pragma unroll
                            for (int u = 0; u < 4; ++u) x[4 * h + u] = cmk(tpk[2 * u], tpk[2 * u + 1]);
                            use(h);
                        }
#endif
                    };
                    if (job == 1) {
#pragma unroll
                        for (int h = 0; h < 8; ++h) {
                            double tpk[8];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int m = 4 * h + u;
                                tpk[2 * u] = x[m].x;
                                tpk[2 * u + 1] = x[m].y;
                                x[m] = nsb_ik(((keepmask >> m) & 1u) ? kxv : 0.0, x[m]);
                            }
                            nsb_tmem_st8(tcol + 16 * h, tpk);
                        }
                        nsb_tmem_wait_st();
                    } else if (job == 2) {
                        unpark([&](int h) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int m = 4 * h + u;
                                x[m] = nsb_ik(((keepmask >> m) & 1u) ? KY[lane + 32 * m] : 0.0, x[m]);
                            }
                        });
                    } else if (job == 3) {
                        unpark([&](int h) {
                            double tpk[8];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int m = 4 * h + u;
                                const double kyv = KY[lane + 32 * m];
                                const double k2 = __dadd_rn(kx2, __dmul_rn(kyv, kyv));
                                double k0 = 0.0, k1 = 0.0;
                                if (k2 > 1e-12) {
                                    const double r = nsb_drcp(k2);
                                    k0 = -(kxv * r);
                                    k1 = -(kyv * r);
                                }
                                if (!((keepmask >> m) & 1u)) k0 = k1 = 0.0;
                                const cd y = nsb_ik(k1, x[m]);
                                tpk[2 * u] = y.x;
                                tpk[2 * u + 1] = y.y;
                                x[m] = nsb_ik(k0, x[m]);
                            }
                            nsb_tmem_st8(tcol + 16 * h, tpk);
                        });
                        nsb_tmem_wait_st();
                    } else {
                        unpark([](int) {}); /* parked ready-made by the third job */
                    }
                } else
#pragma unroll
                for (int m = 0; m < 32; ++m) {
                    const int j1 = lane + 32 * m;
                    const double kyv = KY[j1];
                    /* i k psi = -i (k/k^2) w for f < 2, i k w otherwise: one reciprocal per mode
                     * instead of the reference's two divisions (spectral_core.c:17,66); differs
                     * from it by <= 2 ulp, far inside the 1e-11 parity tolerance */
                    double kk = (f & 1) ? kyv : kxv;
                    if (f < 2) {
                        const double k2 = __dadd_rn(kx2, __dmul_rn(kyv, kyv));
                        kk = (k2 > 1e-12) ? -(kk * nsb_drcp(k2)) : 0.0;
                    }
                    /* 2/3 mask folded into the real factor (bit m of keepmask: row j1 is retained) */
                    kk = ((keepmask >> m) & 1u) ? kk : 0.0;
                    x[m] = nsb_ik(kk, x[m]);
                }
            }
            if (TMA) nsb_fft1024_warp_sw<true>(x, TS, TW, lane);
            else nsb_fft1024_warp_col<true, COLS>(x, S, warp, TW, lane);
        }
        if (job == 0) {
            /* RK update in registers: after the forward transform lane t holds N_hat of rows j1 = t + 32 m of
             * its own column, and the K-layout state of that column is contiguous in j1, so every lane updates
             * its own 32 modes against 512-byte coalesced loads -- no tile round trip, no block barrier, and the
             * new stage state stays in x for the first inverse job. (Round 1 swept the tile with all threads
             * after a barrier: a third of this kernel's stall samples, profiles/README.md round 2.) */
#if NSB_RK_MODE == 1
            if (active) {
                const double dt = *a.rk.dt;
                const bool use_sfd = a.rk.hv.sfd != 0;
#pragma unroll
                for (int m0 = 0; m0 < 32; m0 += NSB_RK_BATCH) {
                    cd W0v[NSB_RK_BATCH], Wsv[NSB_RK_BATCH], ACv[NSB_RK_BATCH], SFv[NSB_RK_BATCH];
#pragma unroll
                    for (int u = 0; u < NSB_RK_BATCH; ++u) {
                        const size_t kid = kbase + lane + 32 * (m0 + u);
                        W0v[u] = a.rk.W0[kid];
                        Wsv[u] = (a.rk.stage == 1) ? W0v[u] : a.rk.Wst[kid];
                        ACv[u] = (a.rk.stage == 1) ? cmk(0.0, 0.0) : a.rk.ACC[kid];
                        SFv[u] = use_sfd ? a.rk.SFD[kid] : cmk(0.0, 0.0);
                    }
#pragma unroll
                    for (int u = 0; u < NSB_RK_BATCH; ++u) {
                        const int m = m0 + u;
                        const int j1 = lane + 32 * m;
                        const size_t kid = kbase + j1;
                        cd nxt = cmk(0.0, 0.0);
                        if ((keepmask >> m) & 1u) {
                            /* x holds the conjugate of the forward transform (forward = conj o inverse o conj) */
                            const cd Nh = cmk(-(x[m].x * g.norm), x[m].y * g.norm);
                            cd acc, sfd;
                            nxt = nsb_rk_math(a.rk, dt, W0v[u], Wsv[u], ACv[u], SFv[u], kxv, KY[j1], ic,
                                              nsb_band(j1 * g.Y2 + j2, g.ny), Nh, &acc, &sfd);
                            if (a.rk.stage < 4) {
                                a.rk.ACC[kid] = acc;
                                a.rk.Wst[kid] = nxt;
                            } else {
                                a.rk.W0[kid] = nxt;
                                if (use_sfd) a.rk.SFD[kid] = sfd;
                            }
                        }
                        x[m] = nxt; /* rows the mask removes get the factor 0 in the inverse jobs */
                    }
                }
                prefetched = true;
            }
#else
            /* The lane parks its 32 values in its own slots of the tile (it reads back only what it wrote
             * itself: no synchronisation) so that the transform registers are free and the update loop can
             * stay rolled: small code, room for the NEXT batch's state loads to be in flight while the current
             * batch is computed. The new stage state goes back into the same slots for the first inverse job. */
            if (active) {
                const double dt = *a.rk.dt;
                const bool use_sfd = a.rk.hv.sfd != 0;
#pragma unroll
                for (int m = 0; m < 32; ++m) S[nsb_sidx<COLS>(lane + 32 * m, warp)] = x[m];
                cd W0v[2][NSB_RK_BATCH], Wsv[2][NSB_RK_BATCH], ACv[2][NSB_RK_BATCH], SFv[2][NSB_RK_BATCH];
                auto fetch = [&](int buf, int m0) {
#pragma unroll
                    for (int u = 0; u < NSB_RK_BATCH; ++u) {
                        const size_t kid = kbase + lane + 32 * (m0 + u);
                        W0v[buf][u] = a.rk.W0[kid];
                        Wsv[buf][u] = (a.rk.stage == 1) ? W0v[buf][u] : a.rk.Wst[kid];
                        ACv[buf][u] = (a.rk.stage == 1) ? cmk(0.0, 0.0) : a.rk.ACC[kid];
                        SFv[buf][u] = use_sfd ? a.rk.SFD[kid] : cmk(0.0, 0.0);
                    }
                };
                auto update = [&](int buf, int m0) {
#pragma unroll
                    for (int u = 0; u < NSB_RK_BATCH; ++u) {
                        const int m = m0 + u;
                        const int j1 = lane + 32 * m;
                        const size_t kid = kbase + j1;
                        cd *slot = &S[nsb_sidx<COLS>(j1, warp)];
                        cd nxt = cmk(0.0, 0.0);
                        if ((keepmask >> m) & 1u) {
                            const cd v = *slot; /* conjugate of the forward transform */
                            const cd Nh = cmk(-(v.x * g.norm), v.y * g.norm);
                            cd acc, sfd;
                            nxt = nsb_rk_math(a.rk, dt, W0v[buf][u], Wsv[buf][u], ACv[buf][u], SFv[buf][u], kxv, KY[j1],
                                              ic, nsb_band(j1 * g.Y2 + j2, g.ny), Nh, &acc, &sfd);
                            if (a.rk.stage < 4) {
                                a.rk.ACC[kid] = acc;
                                a.rk.Wst[kid] = nxt;
                            } else {
                                a.rk.W0[kid] = nxt;
                                if (use_sfd) a.rk.SFD[kid] = sfd;
                            }
                        }
                        if (INV) *slot = nxt; /* rows the mask removes get the factor 0 in the inverse jobs */
                    }
                };
                fetch(0, 0);
#pragma unroll 1
                for (int m0 = 0; m0 < 32; m0 += 2 * NSB_RK_BATCH) {
                    fetch(1, m0 + NSB_RK_BATCH);
                    update(0, m0);
                    if (m0 + 2 * NSB_RK_BATCH < 32) fetch(0, m0 + 2 * NSB_RK_BATCH);
                    update(1, m0 + NSB_RK_BATCH);
                }
                if (INV) {
#pragma unroll
                    for (int m = 0; m < 32; ++m) x[m] = S[nsb_sidx<COLS>(lane + 32 * m, warp)];
                    prefetched = true;
                }
            }
            /* the next transform's transpose (LDG/STG variant: through this column of the tile) writes slots
             * of OTHER lanes: every lane must be done with its parked values first */
            __syncwarp();
#endif
            continue; /* every warp touched only its own column of the tile: no barrier */
        }
        if (TMA && job >= 2) {
            /* the tile store of the previous job ran behind this transform; it must have finished
             * READING the tile before the tile is overwritten */
            if (gt == 0) nsb_tma_store_wait_read();
            NSB_GSYNC();
        }
        if (active) {
            if (PARKT && g.Y2 > 1) {
                NsbTmemRaw8 q[2];
                nsb_tmem_ld8_async(q[0], tcol + 128);
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    double tpk[8];
                    nsb_tmem_wait_ld8(tpk, q[h & 1]);
                    if (h < 7) nsb_tmem_ld8_async(q[(h + 1) & 1], tcol + 128 + 16 * (h + 1));
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int m = 4 * h + u;
                        S[nsb_sidx<COLS>(lane + 32 * m, warp)] = cmulc(x[m], cmk(tpk[2 * u], tpk[2 * u + 1]));
                    }
                }
            } else {
#pragma unroll
            for (int m = 0; m < 32; ++m) {
                cd v = x[m];
                if (g.Y2 > 1) v = cmulc(v, cmul(ta, TB[m]));
                S[nsb_sidx<COLS>(lane + 32 * m, warp)] = v;
            }
            }
            /* registers are free until the next transform: fetch its input now so the
             * latency hides behind the barrier and the cooperative store of this tile */
            if (job < 4 && !PARK) {
#pragma unroll
                for (int m = 0; m < 32; ++m) x[m] = Wre[kbase + lane + 32 * m];
                prefetched = true;
            }
        }
        if (!FWD && job == 4) {
            /* inverse-only launch (stage 1): the first input of the NEXT tile of this CTA, so that its 32 loads per
             * lane are in flight behind this tile's last store instead of stalling the next tile's first transform */
            prefetched = false;
            const int tn = t + GROUPS * (int)gridDim.x;
            if (tn < ntiles) {
                const int j2n = tn / ntx, icn = a.c0 + (tn - j2n * ntx) * COLS + warp;
                if (icn < a.c1) {
                    const size_t kbn = ((size_t)(icn - a.c0) * g.Y2 + j2n) * 1024;
#pragma unroll
                    for (int m = 0; m < 32; ++m) x[m] = Wre[kbn + lane + 32 * m];
                    prefetched = true;
                }
            }
        }
        if (TMA) nsb_fence_proxy_async(); /* generic-proxy tile writes -> visible to the TMA store */
        NSB_GSYNC();
        if (TMA) {
            /* one thread hands the whole tile to the TMA unit: 1024 / tma_br boxes of tma_br rows x COLS
             * columns, each to the device that owns those rows; no barrier, the next transform starts */
            if (gt == 0) {
                for (int r0 = 0; r0 < 1024; r0 += a.tma_br) {
                    const int owner = r0 / a.tma_rl;
                    nsb_tma_3d_s2g(COLS == 4 ? &tm.m4[owner] : &tm.m[owner], S + (size_t)r0 * COLS, 2 * ic0, j2,
                                   (PARK ? ((job + 1) & 3) : job - 1) * a.tma_rl + (r0 - owner * a.tma_rl));
                }
                nsb_tma_store_commit();
            }
            continue;
        } else {
            if (ic0 + cc < a.c1) {
                const size_t fo = (size_t)(job - 1) * a.out_fs + ic0 + cc;
                const int rmask = (1 << a.row_shift) - 1;
#pragma unroll 8
                for (int i = 0; i < 32; ++i) {
                    const int row = rr + 32 * i;
                    const int y = row * g.Y2 + j2;
                    /* the device that owns row y receives it (local store or NVLink peer store) */
                    a.outp[y >> a.row_shift][fo + (size_t)(y & rmask) * a.out_pitch] = S[nsb_sidx<COLS>(row, cc)];
                }
            }
        }
        NSB_GSYNC();
    }
    } /* tile loop */
    if (TMA && INV && gt == 0) nsb_tma_store_wait_read(); /* shared memory must outlive the last store */
    if (PARKT) {
        __syncthreads();
        nsb_tmem_free_all(tbase);
    }
#undef NSB_GSYNC
}

#endif
