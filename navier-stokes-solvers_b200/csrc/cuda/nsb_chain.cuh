/*
 * nsb_chain.cuh -- the physical-space "chain" of one RK stage for X2 == 1024:
 *
 *      T --B--> Z --G--> P --F--> T'
 *
 *   B  (was k_beta_inv)   length-Y2 DFT over j2, Hermitian gather, length-X1 DFT over k1,
 *                         x twiddle, and the Hermitian PACKING of two real fields into one
 *                         complex comb spectrum: Z1 = psi_x + i omega_y, Z2 = psi_y + i omega_x
 *   G  (was k_gamma1024)  per comb: inverse warp FFT-1024 of Z1 and Z2; because of the
 *                         packing the Jacobian is Re*Im of one transform minus Re*Im of the
 *                         other (spectral_core.c:103), so only one real product (8 KB) is
 *                         parked per comb; forward warp FFT-1024 of J(comb 0) + i J(comb 1)
 *   F  (was k_beta_fwd)   split of the packed forward spectrum, x twiddle, length-X1 DFT,
 *                         length-Y2 DFT -> T'
 *
 * The work unit is a SLAB of R = 16 y rows (32 when Y2 == 32; GY = R/Y2 consecutive y1 values): B produces the
 * 16*X1 combs of a slab at once, G consumes them pair by pair, F needs all of them.
 *
 * Two ways to run it:
 *   - three plain kernels (k_chain_B / k_chain_G / k_chain_F), Z and P full size in HBM;
 *   - ONE persistent kernel (k_chain): CTAs draw items from an ordered ticket queue
 *     [F(s-lG-lF) G(s-lG) B(s)] for s = 0.., wait on per-slab completion counters
 *     (release/acquire in global memory) and keep Z and P in small RINGS that stay
 *     resident in the 126 MB L2, so the 4+4+1+1 array passes between B, G and F never
 *     reach HBM. Every item only waits on items with smaller tickets, all CTAs are
 *     co-resident (grid <= SMs x occupancy), hence no deadlock.
 *
 * Replaces, for the hot path: spectral_transforms.c:25-32 (c2r), spectral_core.c:101-111
 * (Jacobian + r2c) and spectral_time_integration.c:14-20 (max |u|).
 */
#ifndef NSB_CHAIN_CUH
#define NSB_CHAIN_CUH

#include "nsb_kernels.cuh"

#define NSB_CH_ROWS 16     /* rows of a slab when Y2 <= 16 */
#define NSB_CH_ROWS_MAX 32 /* Y2 == 32 (32768 rows): one y1 value per slab */

struct ChainArgs {
    const cd *T; /* alpha output, 4 fields x rowsT own rows (pitch PT) */
    int rowsT;   /* y rows this device owns (ny when there is one device) */
    int ybase;   /* first own row */
    cd *Tpp[NSB_MAXP]; /* T' of every device, [ny][cb] own columns: F pushes each column to its owner */
    int cb;      /* columns per device (the pitch of T') */
    cd *Z;       /* [slot][pp][q][1024] packed comb spectra */
    cd *P;       /* [slot][pair][1024] packed forward spectra */
    int nslab;   /* slabs of this launch */
    int slab0;   /* first slab of this launch (slab ranges run concurrently on two streams) */
    int nslotZ, nslotP;
    int R;       /* y rows per slab: a.R, or a.R_MAX when Y2 == 32 */
    int GY;      /* y1 values per slab = R / Y2 */
    int NQ;      /* combs per slab = R * X1 */
    int nbatch;  /* k2 batches */
    int nd, nm;  /* direct / mirror column segments (nsb_beta_segs) */
    unsigned long long *umax_bits;
    /* persistent mode */
    NsbFastDiv fd_perg, fd_nb, fd_nb2; /* TMA-fed B: divisions by nseg*BC, nbatch, 2*nbatch */
    unsigned *ctr; /* [0] ticket, [1 + s] doneB, [1 + nslab + s] doneG, [1 + 2 nslab + s] doneF */
    int leadG, leadF;
};

/* ------------------------------------------------------------------ B items */

template <int Y2>
__device__ __forceinline__ void chain_B_step1(const NsbGeom &g, const ChainArgs &a, cd *tile, int slab, int pp,
                                              int kb) {
    const int nseg = a.nd + a.nm;
    const int fa = pp == 0 ? 0 : 1, fb = pp == 0 ? 3 : 2;
    const int per_g = nseg * g.BC, per_f = a.GY * per_g;
    for (int w = threadIdx.x; w < 2 * per_f; w += blockDim.x) {
        const int fi = w / per_f;
        int r = w - fi * per_f;
        const int gy = r / per_g;
        r -= gy * per_g;
        const int seg = r / g.BC, cc = r - seg * g.BC;
        const int col = nsb_beta_col(g, seg, a.nd, kb, cc, false);
        const int y1 = slab * a.GY + gy;
        cd x[Y2];
        if (col >= 0) {
            const cd *in = a.T + ((size_t)(fi ? fb : fa) * a.rowsT + (size_t)(y1 * Y2)) * g.PT + col;
#pragma unroll
            for (int j2 = 0; j2 < Y2; ++j2) x[j2] = nsb_ldcs(&in[(size_t)j2 * g.PT]);
            nsb_dft<Y2, true>(x);
        } else {
#pragma unroll
            for (int j2 = 0; j2 < Y2; ++j2) x[j2] = cmk(0.0, 0.0);
        }
#pragma unroll
        for (int y2 = 0; y2 < Y2; ++y2)
            tile[(size_t)((fi * a.R + gy * Y2 + y2) * nseg + seg) * g.BC + cc] = x[y2];
    }
}

template <int X1>
__device__ __forceinline__ void chain_gather(cd (&x)[X1], const cd *tile, int row, int nseg, int nd, int nm, int BC,
                                             int cc) {
#pragma unroll
    for (int k1 = 0; k1 < X1; ++k1) {
        if (X1 == 1 || k1 < X1 / 2) {
            x[k1] = (k1 < nd) ? tile[(size_t)(row * nseg + k1) * BC + cc] : cmk(0.0, 0.0);
        } else {
            const int kp = X1 - 1 - k1;
            x[k1] = (kp < nm) ? cconj(tile[(size_t)(row * nseg + nd + kp) * BC + cc]) : cmk(0.0, 0.0);
        }
    }
}

template <int X1>
__device__ __forceinline__ void chain_B_step2(const NsbGeom &g, const ChainArgs &a, const cd *tile, int slot, int pp,
                                              int kb) {
    const int nseg = a.nd + a.nm;
    for (int w = threadIdx.x; w < a.R * g.BC; w += blockDim.x) {
        const int r = w / g.BC, cc = w - r * g.BC;
        const int k2 = kb + cc;
        if (k2 > g.X2h) continue;
        cd A[X1], B[X1];
        chain_gather<X1>(A, tile, r, nseg, a.nd, a.nm, g.BC, cc);
        chain_gather<X1>(B, tile, a.R + r, nseg, a.nd, a.nm, g.BC, cc);
        if (k2 == 0) { /* FFTW c2r ignores Im of the kx = 0 bin */
            A[0].y = 0.0;
            B[0].y = 0.0;
        }
        nsb_dft<X1, true>(A);
        nsb_dft<X1, true>(B);
        cd *zo = a.Z + ((size_t)(slot * 2 + pp) * a.NQ + (size_t)r * X1) * 1024;
#pragma unroll
        for (int x1 = 0; x1 < X1; ++x1) {
            cd av = A[x1], bv = B[x1];
            if (X1 > 1) {
                const cd tw = __ldg(&g.WX[(size_t)k2 * x1]);
                av = cmulc(av, tw);
                bv = cmulc(bv, tw);
            }
            cd *row = zo + (size_t)x1 * 1024;
            /* Z[k] = A[k] + i B[k]; Z[1024-k] = conj(A[k]) + i conj(B[k]) */
            if (k2 == 0 || k2 == g.X2h) {
                row[k2] = cmk(av.x, bv.x);
            } else {
                row[k2] = cmk(av.x - bv.y, av.y + bv.x);
                row[1024 - k2] = cmk(av.x + bv.y, bv.x - av.y);
            }
        }
    }
}

/* item = (slab, pp, batch); tile: 2 * R * nseg * BC complex. All threads call.
 * BIG: also 32-point factors (kept out of the kernels for grids up to 16384: they would cost registers) */
template <bool BIG>
__device__ __forceinline__ void chain_B_item(const NsbGeom &g, const ChainArgs &a, cd *tile, int slab, int slot,
                                             int pp, int batch) {
    const int kb = batch * g.BC;
    switch (g.Y2) {
        case 1: chain_B_step1<1>(g, a, tile, slab, pp, kb); break;
        case 2: chain_B_step1<2>(g, a, tile, slab, pp, kb); break;
        case 4: chain_B_step1<4>(g, a, tile, slab, pp, kb); break;
        case 8: chain_B_step1<8>(g, a, tile, slab, pp, kb); break;
        case 16: chain_B_step1<16>(g, a, tile, slab, pp, kb); break;
        default:
            if constexpr (BIG) chain_B_step1<32>(g, a, tile, slab, pp, kb);
            break;
    }
    __syncthreads();
    switch (g.X1) {
        case 1: chain_B_step2<1>(g, a, tile, slot, pp, kb); break;
        case 2: chain_B_step2<2>(g, a, tile, slot, pp, kb); break;
        case 4: chain_B_step2<4>(g, a, tile, slot, pp, kb); break;
        case 8: chain_B_step2<8>(g, a, tile, slot, pp, kb); break;
        case 16: chain_B_step2<16>(g, a, tile, slot, pp, kb); break;
        default:
            if constexpr (BIG) chain_B_step2<32>(g, a, tile, slot, pp, kb);
            break;
    }
}

/* ------------------------------------------------------------------ G items */

/* per-warp shared memory of a G item: transpose buffer (32x32 doubles, XOR swizzled) |
 * parked products / J of the two combs (2 x 1024 doubles) | psi_x of the current comb (UMAX) */
#define NSB_CHG_WARP_BYTES(UMAX) (8192 + 16384 + ((UMAX) ? 8192 : 0))
#define NSB_CH_TW_BYTES (1024 * 16)

/* one comb pair by one warp. wb: this warp's NSB_CHG_WARP_BYTES; returns max |u|^2 seen */
template <bool UMAX>
__device__ __forceinline__ double chain_G_pair(const cd *z1base, const cd *z2base, cd *pout, unsigned char *wb,
                                               const cd *TW, int lane, const cd *znext = nullptr) {
    double *T = reinterpret_cast<double *>(wb);
    double *PK = reinterpret_cast<double *>(wb + 8192);
    double *PX = reinterpret_cast<double *>(wb + 8192 + 16384);
    double umax = 0.0;
    cd x[32];
#pragma unroll 1
    for (int comb = 0; comb < 2; ++comb) {
        const cd *z1 = z1base + (size_t)comb * 1024, *z2 = z2base + (size_t)comb * 1024;
#pragma unroll
        for (int m = 0; m < 32; ++m) x[m] = nsb_ldcg(&z1[lane + 32 * m]);
        /* pull the next 16 KB input towards L2 while this transform runs */
#pragma unroll
        for (int i = 0; i < 4; ++i) nsb_prefetch_l2(reinterpret_cast<const char *>(z2) + (lane * 4 + i) * 128);
        nsb_fft1024_warp_sw<true>(x, T, TW, lane); /* x = psi_x + i omega_y */
#pragma unroll
        for (int m = 0; m < 32; ++m) {
            PK[comb * 1024 + m * 32 + lane] = __dmul_rn(x[m].x, x[m].y);
            if (UMAX) PX[m * 32 + lane] = x[m].x;
        }
#pragma unroll
        for (int m = 0; m < 32; ++m) x[m] = nsb_ldcg(&z2[lane + 32 * m]);
        if (comb == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                nsb_prefetch_l2(reinterpret_cast<const char *>(z1base + 1024) + (lane * 4 + i) * 128);
        }
        nsb_fft1024_warp_sw<true>(x, T, TW, lane); /* x = psi_y + i omega_x */
#pragma unroll
        for (int m = 0; m < 32; ++m) {
            /* spectral_core.c:103 : J = psi_x*omega_y - psi_y*omega_x */
            const double J = __dsub_rn(PK[comb * 1024 + m * 32 + lane], __dmul_rn(x[m].x, x[m].y));
            if (UMAX) { /* spectral_time_integration.c:16 : |u|^2 = psi_y^2 + psi_x^2 */
                const double px = PX[m * 32 + lane];
                umax = nsb_max_sticky(umax, __dadd_rn(__dmul_rn(x[m].x, x[m].x), __dmul_rn(px, px)));
            }
            if (comb == 0) PK[m * 32 + lane] = J;
            else x[m] = cmk(PK[m * 32 + lane], J);
        }
    }
    if (znext != nullptr) { /* first input of this warp's NEXT pair (persistent kernel) */
#pragma unroll
        for (int i = 0; i < 4; ++i) nsb_prefetch_l2(reinterpret_cast<const char *>(znext) + (lane * 4 + i) * 128);
    }
    nsb_fft1024_warp_sw<false>(x, T, TW, lane);
#pragma unroll
    for (int m = 0; m < 32; ++m) pout[lane + 32 * m] = x[m];
    return umax;
}

/* item = (slab, group of W pairs). ub: W * NSB_CHG_WARP_BYTES. */
template <bool UMAX>
__device__ __forceinline__ void chain_G_item(const ChainArgs &a, unsigned char *ub, const cd *TW, int slotZ, int slotP,
                                             int group) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int NP = a.NQ >> 1;
    const int pair = group * W + warp;
    if (pair >= NP) return;
    const cd *z1 = a.Z + ((size_t)(slotZ * 2 + 0) * a.NQ + 2 * pair) * 1024;
    const cd *z2 = a.Z + ((size_t)(slotZ * 2 + 1) * a.NQ + 2 * pair) * 1024;
    cd *po = a.P + ((size_t)slotP * NP + pair) * 1024;
    double umax = chain_G_pair<UMAX>(z1, z2, po, ub + (size_t)warp * NSB_CHG_WARP_BYTES(UMAX), TW, lane);
    if (UMAX) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) umax = nsb_max_sticky(umax, __shfl_xor_sync(0xffffffffu, umax, off));
        if (lane == 0) atomicMax(a.umax_bits, (unsigned long long)__double_as_longlong(umax));
    }
}

/* ------------------------------------------------------------------ F items */

template <int X1>
__device__ __forceinline__ void chain_F_step1(const NsbGeom &g, const ChainArgs &a, cd *tile, int slot, int kb,
                                              const cd *twtab) {
    const int nseg = a.nd + a.nm;
    const int NP = a.NQ >> 1;
    for (int w = threadIdx.x; w < a.R * g.BC; w += blockDim.x) {
        const int r = w / g.BC, cc = w - r * g.BC;
        const int k2 = kb + cc;
        cd x[X1];
        if (k2 <= g.X2h) {
            const int km = (1024 - k2) & 1023;
            if (X1 == 1) {
                const cd *pr = a.P + ((size_t)slot * NP + (r >> 1)) * 1024;
                const cd Pk = nsb_ldcg(&pr[k2]), Q = nsb_ldcg(&pr[km]);
                x[0] = (r & 1) ? cmk(0.5 * (Pk.y + Q.y), -0.5 * (Pk.x - Q.x)) : cmk(0.5 * (Pk.x + Q.x), 0.5 * (Pk.y - Q.y));
            } else {
#pragma unroll
                for (int i = 0; i < X1 / 2; ++i) {
                    const cd *pr = a.P + ((size_t)slot * NP + (size_t)(r * X1) / 2 + i) * 1024;
                    const cd Pk = nsb_ldcg(&pr[k2]), Q = nsb_ldcg(&pr[km]);
                    const cd e = cmk(0.5 * (Pk.x + Q.x), 0.5 * (Pk.y - Q.y));
                    const cd o = cmk(0.5 * (Pk.y + Q.y), -0.5 * (Pk.x - Q.x));
                    x[2 * i] = cmul(e, twtab ? twtab[cc * X1 + 2 * i] : __ldg(&g.WX[(size_t)k2 * (2 * i)]));
                    x[2 * i + 1] = cmul(o, twtab ? twtab[cc * X1 + 2 * i + 1] : __ldg(&g.WX[(size_t)k2 * (2 * i + 1)]));
                }
            }
            nsb_dft<X1, false>(x);
        } else {
#pragma unroll
            for (int x1 = 0; x1 < X1; ++x1) x[x1] = cmk(0.0, 0.0);
        }
#pragma unroll
        for (int m = 0; m < X1; ++m) {
            if (X1 == 1 || m < X1 / 2) {
                if (m < a.nd) tile[(size_t)(r * nseg + m) * g.BC + cc] = x[m];
            } else {
                const int kp = X1 - 1 - m;
                if (kp < a.nm) tile[(size_t)(r * nseg + a.nd + kp) * g.BC + cc] = cconj(x[m]);
            }
        }
    }
}

template <int Y2>
__device__ __forceinline__ void chain_F_step2(const NsbGeom &g, const ChainArgs &a, const cd *tile, int slab, int kb) {
    const int nseg = a.nd + a.nm;
    const int per_g = nseg * g.BC;
    for (int w = threadIdx.x; w < a.GY * per_g; w += blockDim.x) {
        const int gy = w / per_g;
        const int r = w - gy * per_g;
        const int seg = r / g.BC, cc = r - seg * g.BC;
        const int col = nsb_beta_col(g, seg, a.nd, kb, cc, true);
        if (col < 0) continue;
        cd x[Y2];
#pragma unroll
        for (int y2 = 0; y2 < Y2; ++y2) x[y2] = tile[(size_t)((gy * Y2 + y2) * nseg + seg) * g.BC + cc];
        nsb_dft<Y2, false>(x);
        /* the device that owns column `col` receives it (local store or NVLink peer store) */
        const int owner = col / a.cb;
        cd *out = a.Tpp[owner] + (size_t)(a.ybase + (slab * a.GY + gy) * Y2) * a.cb + (col - owner * a.cb);
#pragma unroll
        for (int j2 = 0; j2 < Y2; ++j2) nsb_stcs(&out[(size_t)j2 * a.cb], x[j2]);
    }
}

/* item = (slab, batch); tile: R * nseg * BC complex. All threads call. */
template <bool BIG>
__device__ __forceinline__ void chain_F_item(const NsbGeom &g, const ChainArgs &a, cd *tile, int slab, int slot,
                                             int batch) {
    const int kb = batch * g.BC;
    /* measured: staging the x twiddles in shared memory (as B does) costs F more in the extra
     * barrier than the L1/L2 twiddle loads do */
    const cd *twtab = nullptr;
    switch (g.X1) {
        case 1: chain_F_step1<1>(g, a, tile, slot, kb, twtab); break;
        case 2: chain_F_step1<2>(g, a, tile, slot, kb, twtab); break;
        case 4: chain_F_step1<4>(g, a, tile, slot, kb, twtab); break;
        case 8: chain_F_step1<8>(g, a, tile, slot, kb, twtab); break;
        case 16: chain_F_step1<16>(g, a, tile, slot, kb, twtab); break;
        default:
            if constexpr (BIG) chain_F_step1<32>(g, a, tile, slot, kb, twtab);
            break;
    }
    __syncthreads();
    switch (g.Y2) {
        case 1: chain_F_step2<1>(g, a, tile, slab, kb); break;
        case 2: chain_F_step2<2>(g, a, tile, slab, kb); break;
        case 4: chain_F_step2<4>(g, a, tile, slab, kb); break;
        case 8: chain_F_step2<8>(g, a, tile, slab, kb); break;
        case 16: chain_F_step2<16>(g, a, tile, slab, kb); break;
        default:
            if constexpr (BIG) chain_F_step2<32>(g, a, tile, slab, kb);
            break;
    }
}

/* ------------------------------------------------ B, TMA-fed and persistent */
/* requires R*BC*16 bytes per box row block, i.e. BC*2 doubles <= 256 and R <= 256 rows */

/*
 * Valid part of column segment `seg` for the k2 batch starting at kb, in ASCENDING memory
 * order: `cnt` complex values starting at column `col0`, landing at tile positions
 * pos0 .. pos0+cnt-1. Direct segments hold k2 = kb + pos, mirror segments hold
 * k2 = kb + (BC-1-pos) (their columns run backwards in k2).
 */
__device__ __forceinline__ void chain_seg_range(const NsbGeom &g, int nd, int seg, int kb, int *pos0, int *cnt,
                                                int *col0) {
    if (seg < nd) {
        int n = g.BC;
        if (g.X2h + 1 - kb < n) n = g.X2h + 1 - kb;
        if (g.imax + 1 - seg * g.X2 - kb < n) n = g.imax + 1 - seg * g.X2 - kb;
        *pos0 = 0;
        *cnt = n > 0 ? n : 0;
        *col0 = seg * g.X2 + kb;
    } else {
        const int base = (seg - nd) * g.X2 + g.X2 - kb; /* column of cc = 0 */
        int cmin = base - g.imax;
        if (cmin < 0) cmin = 0;
        int cmax = g.BC - 1;
        if (g.X2h - kb < cmax) cmax = g.X2h - kb;
        *pos0 = g.BC - 1 - cmax;
        *cnt = cmax >= cmin ? cmax - cmin + 1 : 0;
        *col0 = base - cmax;
    }
}

/*
 * TMA tensor copies of the raw input of item (slab, pp, batch): one 16-row x BC-column box
 * (4 KB) per (field, segment) through the tensor map of T (rows = 4*ny, pitch PT), landing
 * at tile[(fi*nseg + seg)*16 + row][pos]. Columns outside the retained set come back as
 * zeros or as neighbouring retained columns; step 1 masks them by range. Completion on `bar`.
 */
#ifndef NSB_B_TW_TMA
#define NSB_B_TW_TMA 1
#endif
template <bool TW_TMA>
__device__ __forceinline__ void chain_B_issue(const NsbGeom &g, const ChainArgs &a, const nsb_tmap *tmT, cd *tile,
                                              cd *twtab, nsb_mbar_t *bar, int slab, int pp, int kb) {
    const int nseg = a.nd + a.nm;
    if (threadIdx.x == 0) {
        nsb_fence_proxy_async();
        const unsigned twb = TW_TMA ? (unsigned)(g.BC * NSB_TWK_PITCH(g.X1) * 16) : 0u;
        nsb_mbar_expect_tx(bar, (unsigned)(2 * nseg) * (unsigned)(a.R * g.BC * 16) + twb);
        /* the x twiddles of this k2 batch: one contiguous block of the batch-ordered table */
        if (TW_TMA) nsb_bulk_g2s(twtab, g.TWK + (size_t)kb * NSB_TWK_PITCH(g.X1), twb, bar);
    }
    __syncthreads();
    if ((int)threadIdx.x < 2 * nseg) {
        const int fi = threadIdx.x / nseg, seg = threadIdx.x - fi * nseg;
        const int f = pp == 0 ? (fi ? 3 : 0) : (fi ? 2 : 1);
        const int col0 = (seg < a.nd) ? seg * g.X2 + kb : (seg - a.nd) * g.X2 + g.X2 - kb - (g.BC - 1);
        nsb_tma_2d_g2s(tile + (size_t)(fi * nseg + seg) * a.R * g.BC, tmT, 2 * col0,
                       f * a.rowsT + slab * a.R, bar);
    }
}

/* step 1 on a TMA-landed tile, in place: positions outside the valid range become zeros */
template <int Y2>
__device__ __forceinline__ void chain_B2_step1(const NsbGeom &g, const ChainArgs &a, cd *tile, int kb) {
    const int nseg = a.nd + a.nm;
    const int per_g = nseg * g.BC, per_f = a.GY * per_g;
    const int bcs = nsb_log2i(g.BC);
    for (int w = threadIdx.x; w < 2 * per_f; w += blockDim.x) {
        const int fi = w >= per_f ? 1 : 0;
        int r = w - fi * per_f;
        const int gy = (int)nsb_fastdiv((unsigned)r, a.fd_perg);
        r -= gy * per_g;
        const int seg = r >> bcs, pos = r & (g.BC - 1);
        int p0, cnt, c0;
        chain_seg_range(g, a.nd, seg, kb, &p0, &cnt, &c0);
        cd *col = tile + (size_t)((fi * nseg + seg) * a.R + gy * Y2) * g.BC + pos;
        cd x[Y2];
        if (pos >= p0 && pos < p0 + cnt) {
#pragma unroll
            for (int j2 = 0; j2 < Y2; ++j2) x[j2] = col[j2 * g.BC];
            nsb_dft<Y2, true>(x);
        } else {
#pragma unroll
            for (int j2 = 0; j2 < Y2; ++j2) x[j2] = cmk(0.0, 0.0);
        }
#pragma unroll
        for (int y2 = 0; y2 < Y2; ++y2) col[y2 * g.BC] = x[y2];
    }
}

/* gather for tiles in memory order (mirror segments reversed) */
template <int X1>
__device__ __forceinline__ void chain_gather_mo(cd (&x)[X1], const cd *tile, int fi, int r, int nseg, int nd, int nm,
                                                int BC, int cc, int R) {
#pragma unroll
    for (int k1 = 0; k1 < X1; ++k1) {
        if (X1 == 1 || k1 < X1 / 2) {
            x[k1] = (k1 < nd) ? tile[(size_t)((fi * nseg + k1) * R + r) * BC + cc] : cmk(0.0, 0.0);
        } else {
            const int kp = X1 - 1 - k1;
            x[k1] = (kp < nm) ? cconj(tile[(size_t)((fi * nseg + nd + kp) * R + r) * BC + (BC - 1 - cc)])
                              : cmk(0.0, 0.0);
        }
    }
}

/* twtab (nullable): shared-memory copy of the x twiddles of this batch, twtab[cc*(X1+1) + x1]: rows padded by one
 * element so that the 8 lanes of a 128-bit shared-memory phase (consecutive cc, same x1) fall into 8 different
 * 16-byte bank groups (with pitch X1 they were an 8-way conflict: 60 % of the round-1 kernel's shared-memory
 * wavefronts and a 32-cycle pipe occupancy per twiddle load, which the twiddle multiply then waited for). */
template <int X1, bool MO>
__device__ __forceinline__ void chain_B_step2x(const NsbGeom &g, const ChainArgs &a, const cd *tile, int slot, int pp,
                                               int kb, const cd *twtab) {
    const int nseg = a.nd + a.nm;
    const int bcs = nsb_log2i(g.BC);
    for (int w = threadIdx.x; w < a.R * g.BC; w += blockDim.x) {
        const int r = w >> bcs, cc = w & (g.BC - 1);
        const int k2 = kb + cc;
        if (k2 > g.X2h) continue;
        cd A[X1], B[X1];
        if (MO) {
            chain_gather_mo<X1>(A, tile, 0, r, nseg, a.nd, a.nm, g.BC, cc, a.R);
            chain_gather_mo<X1>(B, tile, 1, r, nseg, a.nd, a.nm, g.BC, cc, a.R);
        } else {
            chain_gather<X1>(A, tile, r, nseg, a.nd, a.nm, g.BC, cc);
            chain_gather<X1>(B, tile, a.R + r, nseg, a.nd, a.nm, g.BC, cc);
        }
        if (k2 == 0) { /* FFTW c2r ignores Im of the kx = 0 bin */
            A[0].y = 0.0;
            B[0].y = 0.0;
        }
        nsb_dft<X1, true>(A);
        nsb_dft<X1, true>(B);
        cd *zo = a.Z + ((size_t)(slot * 2 + pp) * a.NQ + (size_t)r * X1) * 1024;
#pragma unroll
        for (int x1 = 0; x1 < X1; ++x1) {
            cd av = A[x1], bv = B[x1];
            if (X1 > 1) {
                const cd tw = twtab ? twtab[cc * NSB_TWK_PITCH(X1) + x1] : __ldg(&g.WX[(size_t)k2 * x1]);
                av = cmulc(av, tw);
                bv = cmulc(bv, tw);
            }
            cd *row = zo + (size_t)x1 * 1024;
            /* Z[k] = A[k] + i B[k]; Z[1024-k] = conj(A[k]) + i conj(B[k]) */
            if (k2 == 0 || k2 == g.X2h) {
                row[k2] = cmk(av.x, bv.x);
            } else {
                row[k2] = cmk(av.x - bv.y, av.y + bv.x);
                row[1024 - k2] = cmk(av.x + bv.y, bv.x - av.y);
            }
        }
    }
}

/*
 * Step 2 for 32-point factors (X1 == 32), tiles in memory order. Two 32-point register DFTs per thread
 * (A and B, as chain_B_step2x does) would need 256 registers for the data alone, and a 32-row x 4-column
 * batch has only 128 (row, k2) items for 256 threads. The DFT is linear, so the packing moves IN FRONT of
 * it: one thread transforms C = A + iB (-> Z[k2]), its neighbour D = A - iB (-> Z[1024-k2] = conj(D)):
 * 256 items, one 32-point array each. At the two self-conjugate bins (k2 = 0 and 512) the stored value
 * (Re a, Re b) needs both halves: ((C.x + D.x)/2, (C.y - D.y)/2), exchanged by one shuffle with lane^1.
 * blockDim.x and 2*R*BC are multiples of 32, so a pair never straddles a warp and all lanes shuffle.
 */
template <int X1>
__device__ __forceinline__ void chain_B_step2p(const NsbGeom &g, const ChainArgs &a, const cd *tile, int slot, int pp,
                                               int kb, const cd *twtab) {
    const int nseg = a.nd + a.nm;
    const bool special_batch = (kb == 0) || (kb <= g.X2h && kb + g.BC > g.X2h);
    for (int w = threadIdx.x; w < 2 * a.R * g.BC; w += blockDim.x) {
        /* the +- halves of an item sit in neighbouring lanes only where they must exchange (the two
         * batches with a self-conjugate bin); elsewhere each half takes a contiguous run of items, so the
         * 8 lanes of a 128-bit shared-memory phase read 8 distinct values instead of 4 pairs */
        const int nitem = a.R * g.BC;
        const int half = special_batch ? (w & 1) : (w >= nitem ? 1 : 0);
        const int item = special_batch ? (w >> 1) : (w - half * nitem);
        const int r = item >> nsb_log2i(g.BC), cc = item & (g.BC - 1);
        const int k2 = kb + cc;
        const bool valid = k2 <= g.X2h;
        const double sg = half ? -1.0 : 1.0;
        cd x[X1];
#pragma unroll
        for (int k1 = 0; k1 < X1; ++k1) {
            cd av, bv;
            if (k1 < X1 / 2) {
                if (k1 < a.nd) {
                    av = tile[(size_t)((0 * nseg + k1) * a.R + r) * g.BC + cc];
                    bv = tile[(size_t)((1 * nseg + k1) * a.R + r) * g.BC + cc];
                } else {
                    av = bv = cmk(0.0, 0.0);
                }
            } else {
                const int kp = X1 - 1 - k1;
                if (kp < a.nm) {
                    av = cconj(tile[(size_t)((0 * nseg + a.nd + kp) * a.R + r) * g.BC + (g.BC - 1 - cc)]);
                    bv = cconj(tile[(size_t)((1 * nseg + a.nd + kp) * a.R + r) * g.BC + (g.BC - 1 - cc)]);
                } else {
                    av = bv = cmk(0.0, 0.0);
                }
            }
            if (k1 == 0 && k2 == 0) { /* FFTW c2r ignores Im of the kx = 0 bin */
                av.y = 0.0;
                bv.y = 0.0;
            }
            x[k1] = cmk(av.x - sg * bv.y, av.y + sg * bv.x); /* A + iB  |  A - iB */
        }
        nsb_dft<X1, true>(x);
        cd *zo = a.Z + ((size_t)(slot * 2 + pp) * a.NQ + (size_t)r * X1) * 1024;
        const bool special = (k2 == 0 || k2 == g.X2h);
#pragma unroll
        for (int x1 = 0; x1 < X1; ++x1) {
            cd v = cmulc(x[x1], twtab[cc * NSB_TWK_PITCH(X1) + x1]);
            if (special_batch) { /* batch-uniform: every lane shuffles */
                const double ox = __shfl_xor_sync(0xffffffffu, v.x, 1), oy = __shfl_xor_sync(0xffffffffu, v.y, 1);
                if (special) v = cmk(0.5 * (v.x + ox), 0.5 * (v.y - oy)); /* half 0: (Re a, Re b) */
            }
            if (!valid) continue;
            cd *row = zo + (size_t)x1 * 1024;
            if (half == 0) row[k2] = v;
            else if (!special) row[1024 - k2] = cmk(v.x, -v.y);
        }
    }
}

/*
 * Persistent, TMA-fed B: grid = SMs, each CTA walks items blockIdx.x + n*gridDim.x
 * (item = (slab, pp, batch), batch fastest). Two shared-memory stages: while the CTA
 * transforms the tile of item n, the bulk copies of item n+1 are in flight, so HBM latency
 * is off the critical path with one CTA per SM.
 * smem: 2 stages x (2*16*nseg*BC complex) + 2 mbarriers (+128 B alignment slack).
 */
/* SPLIT: step 2 with the Hermitian packing in front of the k1 DFT (chain_B_step2p: one thread transforms A + iB, another
 * A - iB, ONE 16-point array per thread instead of two): half the registers per thread, so 512 threads fit one SM. */
template <int NT, int MINB, bool BIG, bool SPLIT = false>
__global__ void __launch_bounds__(NT, MINB) k_chain_B_tma(NsbGeom g, ChainArgs a, const NSB_GRID_CONSTANT nsb_tmap tmT) {
    NSB_DYN_SMEM(cd, smem_raw);
    /* 128-byte alignment for the TMA destination. (Measured: keeping shared-memory provenance, i.e.
     * LDS/STS instead of the generic LD/ST this cast produces, made the kernel 7 % SLOWER.) */
    cd *smem = reinterpret_cast<cd *>((reinterpret_cast<size_t>(smem_raw) + 127) & ~(size_t)127);
    const int nseg = a.nd + a.nm;
    const size_t tile_elems = (size_t)2 * a.R * nseg * g.BC;
    nsb_mbar_t *bars = reinterpret_cast<nsb_mbar_t *>(smem + 2 * tile_elems);
    cd *twtabs = smem + 2 * tile_elems + 4; /* 2 stages x BC * X1 x twiddles (TMA bulk copies, with the tile) */
    const size_t tw_elems = (size_t)g.BC * NSB_TWK_PITCH(g.X1);
    if (threadIdx.x == 0) {
        nsb_mbar_init(&bars[0], 1);
        nsb_mbar_init(&bars[1], 1);
    }
    __syncthreads();
    const int nitems = a.nslab * 2 * a.nbatch;
    unsigned phase[2] = {0u, 0u};
    int item = blockIdx.x;
    auto decode = [&](int it, int *slab, int *pp, int *kb) {
        const int sl = (int)nsb_fastdiv((unsigned)it, a.fd_nb2), rem = it - sl * 2 * a.nbatch;
        const int p = (int)nsb_fastdiv((unsigned)rem, a.fd_nb);
        *slab = a.slab0 + sl;
        *pp = p;
        *kb = (rem - p * a.nbatch) * g.BC;
    };
    if (item < nitems) {
        int slab, pp, kb;
        decode(item, &slab, &pp, &kb);
        chain_B_issue<NSB_B_TW_TMA && !BIG>(g, a, &tmT, smem, twtabs, &bars[0], slab, pp, kb);
    }
    for (int n = 0; item < nitems; ++n, item += gridDim.x) {
        const int st = n & 1;
        const int next = item + gridDim.x;
        if (next < nitems) {
            int slab, pp, kb;
            decode(next, &slab, &pp, &kb);
            chain_B_issue<NSB_B_TW_TMA && !BIG>(g, a, &tmT, smem + (size_t)(st ^ 1) * tile_elems, twtabs + (size_t)(st ^ 1) * tw_elems,
                               &bars[st ^ 1], slab, pp, kb);
        }
        nsb_mbar_wait(&bars[st], phase[st]);
        phase[st] ^= 1u;
        cd *tile = smem + (size_t)st * tile_elems;
        int slab, pp, kb;
        decode(item, &slab, &pp, &kb);
        if constexpr (BIG || !NSB_B_TW_TMA) { /* 32-point factors: the twiddles of the batch are gathered by the threads
                              * (measured: the TMA-staged table costs this variant a factor of two) */
            cd *tw = twtabs + (size_t)st * tw_elems;
            for (int w = threadIdx.x; w < g.BC * g.X1; w += blockDim.x) {
                const int cc = w / g.X1, x1 = w - cc * g.X1;
                tw[cc * NSB_TWK_PITCH(g.X1) + x1] = (kb + cc <= g.X2h) ? __ldg(&g.WX[(size_t)(kb + cc) * x1]) : cmk(1.0, 0.0);
            }
        }
        switch (g.Y2) {
            case 1: chain_B2_step1<1>(g, a, tile, kb); break;
            case 2: chain_B2_step1<2>(g, a, tile, kb); break;
            case 4: chain_B2_step1<4>(g, a, tile, kb); break;
            case 8: chain_B2_step1<8>(g, a, tile, kb); break;
            case 16: chain_B2_step1<16>(g, a, tile, kb); break;
            default:
                if constexpr (BIG) chain_B2_step1<32>(g, a, tile, kb);
                break;
        }
        __syncthreads();
        const cd *twtab = twtabs + (size_t)st * tw_elems;
        const int slot = slab; /* plain chain: every slab has its own Z slot (nslotZ == slabs of the device) */
        if constexpr (SPLIT) {
            switch (g.X1) {
                case 2: chain_B_step2p<2>(g, a, tile, slot, pp, kb, twtab); break;
                case 4: chain_B_step2p<4>(g, a, tile, slot, pp, kb, twtab); break;
                case 8: chain_B_step2p<8>(g, a, tile, slot, pp, kb, twtab); break;
                case 16: chain_B_step2p<16>(g, a, tile, slot, pp, kb, twtab); break;
                default: break;
            }
        } else {
            switch (g.X1) {
                case 1: chain_B_step2x<1, true>(g, a, tile, slot, pp, kb, twtab); break;
                case 2: chain_B_step2x<2, true>(g, a, tile, slot, pp, kb, twtab); break;
                case 4: chain_B_step2x<4, true>(g, a, tile, slot, pp, kb, twtab); break;
                case 8: chain_B_step2x<8, true>(g, a, tile, slot, pp, kb, twtab); break;
                case 16: chain_B_step2x<16, true>(g, a, tile, slot, pp, kb, twtab); break;
                default:
                    if constexpr (BIG) chain_B_step2p<32>(g, a, tile, slot, pp, kb, twtab);
                    break;
            }
        }
        __syncthreads(); /* the tile is free for the copies of item n+2 */
    }
}


/* ------------------------------------------------ F, TMA-fed and persistent */

/*
 * Step 1 of F on TMA-landed boxes: A[row][cc] = P[pair row][kb + cc], M[row][cc'] = P[pair row][1024 - kb - (BC-1) + cc']
 * (the mirror bins 1024 - k2 in ascending memory order, so bin 1024 - (kb + cc) sits at cc' = BC - 1 - cc), rows =
 * the NP = R * X1 / 2 comb pairs of the slab. Same arithmetic as chain_F_step1 (results agree to rounding: the two
 * kernels are contracted into FMAs differently by the compiler).
 */
template <int X1>
__device__ __forceinline__ void chain_F2_step1(const NsbGeom &g, const ChainArgs &a, cd *tile, const cd *A, const cd *M,
                                               const cd *twtab, int kb) {
    const int nseg = a.nd + a.nm;
    const int bcs = nsb_log2i(g.BC);
    for (int w = threadIdx.x; w < a.R * g.BC; w += blockDim.x) {
        const int r = w >> bcs, cc = w & (g.BC - 1);
        const int k2 = kb + cc;
        cd x[X1];
        if (k2 <= g.X2h) {
#pragma unroll
            for (int i = 0; i < X1 / 2; ++i) {
                const int row = (r * X1) / 2 + i;
                const cd Pk = A[row * g.BC + cc];
                const cd Q = (k2 == 0) ? Pk : M[row * g.BC + (g.BC - 1 - cc)]; /* bin 1024 wraps to bin 0 */
                const cd e = cmk(0.5 * (Pk.x + Q.x), 0.5 * (Pk.y - Q.y));
                const cd o = cmk(0.5 * (Pk.y + Q.y), -0.5 * (Pk.x - Q.x));
                x[2 * i] = cmul(e, twtab[cc * NSB_TWK_PITCH(X1) + 2 * i]);
                x[2 * i + 1] = cmul(o, twtab[cc * NSB_TWK_PITCH(X1) + 2 * i + 1]);
            }
            nsb_dft<X1, false>(x);
        } else {
#pragma unroll
            for (int x1 = 0; x1 < X1; ++x1) x[x1] = cmk(0.0, 0.0);
        }
#pragma unroll
        for (int m = 0; m < X1; ++m) {
            if (m < X1 / 2) {
                if (m < a.nd) tile[(size_t)(r * nseg + m) * g.BC + cc] = x[m];
            } else {
                const int kp = X1 - 1 - m;
                if (kp < a.nm) tile[(size_t)(r * nseg + a.nd + kp) * g.BC + cc] = cconj(x[m]);
            }
        }
    }
}

/*
 * Persistent F (grids with 2 <= X1 <= 16): grid = SMs, each CTA walks items blockIdx.x + n * gridDim.x (item =
 * (slab, k2 batch), batch fastest). The two P boxes of item n+1 (the k2 batch and its mirror bins, NP rows x BC
 * columns each) are fetched by tensor-map TMA while item n is transformed: round 1's F spent 63 % of its stall
 * samples waiting for the 16 scattered P loads per thread at the top of every CTA.
 * smem: 2 stages x 2 boxes (NP * BC complex each) | tile (R * nseg * BC complex) | 2 mbarriers.
 */
__global__ void __launch_bounds__(256, 1) k_chain_F_tma(NsbGeom g, ChainArgs a, const NSB_GRID_CONSTANT nsb_tmap tmP) {
    NSB_DYN_SMEM(cd, smem_raw);
    cd *smem = reinterpret_cast<cd *>((reinterpret_cast<size_t>(smem_raw) + 127) & ~(size_t)127);
    const int nseg = a.nd + a.nm;
    const int NP = a.NQ >> 1;
    const size_t box = (size_t)NP * g.BC;
    const size_t tw_elems = (size_t)g.BC * NSB_TWK_PITCH(g.X1);
    cd *tile = smem + 4 * box;
    cd *twtabs = tile + (size_t)a.R * nseg * g.BC; /* 2 stages of x twiddles (the B kernel's batch-ordered table) */
    nsb_mbar_t *bars = reinterpret_cast<nsb_mbar_t *>(twtabs + 2 * tw_elems);
    if (threadIdx.x == 0) {
        nsb_mbar_init(&bars[0], 1);
        nsb_mbar_init(&bars[1], 1);
    }
    __syncthreads();
    const int nitems = a.nslab * a.nbatch;
    auto issue = [&](int it, int st) {
        if (threadIdx.x == 0) {
            const int sl = (int)nsb_fastdiv((unsigned)it, a.fd_nb), kb = (it - sl * a.nbatch) * g.BC;
            const int row0 = (a.slab0 + sl) * NP;
            cd *dst = smem + (size_t)st * 2 * box;
            nsb_fence_proxy_async();
            nsb_mbar_expect_tx(&bars[st], (unsigned)(2 * box * 16) + (unsigned)(tw_elems * 16));
            nsb_bulk_g2s(twtabs + (size_t)st * tw_elems, g.TWK + (size_t)kb * NSB_TWK_PITCH(g.X1), (unsigned)(tw_elems * 16),
                         &bars[st]);
            nsb_tma_2d_g2s(dst, &tmP, 2 * kb, row0, &bars[st]);
            nsb_tma_2d_g2s(dst + box, &tmP, 2 * (1024 - kb - (g.BC - 1)), row0, &bars[st]);
        }
    };
    unsigned phase[2] = {0u, 0u};
    int item = blockIdx.x;
    if (item < nitems) issue(item, 0);
    for (int n = 0; item < nitems; ++n, item += gridDim.x) {
        const int st = n & 1;
        if (item + (int)gridDim.x < nitems) issue(item + gridDim.x, st ^ 1); /* stage st^1 was released by the barrier below */
        nsb_mbar_wait(&bars[st], phase[st]);
        phase[st] ^= 1u;
        const int sl = (int)nsb_fastdiv((unsigned)item, a.fd_nb), kb = (item - sl * a.nbatch) * g.BC;
        const cd *A = smem + (size_t)st * 2 * box, *M = A + box, *tw = twtabs + (size_t)st * tw_elems;
        switch (g.X1) {
            case 2: chain_F2_step1<2>(g, a, tile, A, M, tw, kb); break;
            case 4: chain_F2_step1<4>(g, a, tile, A, M, tw, kb); break;
            case 8: chain_F2_step1<8>(g, a, tile, A, M, tw, kb); break;
            case 16: chain_F2_step1<16>(g, a, tile, A, M, tw, kb); break;
            default: break;
        }
        __syncthreads(); /* tile complete; the boxes of this stage are free for item n+2 */
        switch (g.Y2) {
            case 1: chain_F_step2<1>(g, a, tile, a.slab0 + sl, kb); break;
            case 2: chain_F_step2<2>(g, a, tile, a.slab0 + sl, kb); break;
            case 4: chain_F_step2<4>(g, a, tile, a.slab0 + sl, kb); break;
            case 8: chain_F_step2<8>(g, a, tile, a.slab0 + sl, kb); break;
            case 16: chain_F_step2<16>(g, a, tile, a.slab0 + sl, kb); break;
            default: break;
        }
        __syncthreads(); /* the tile is rewritten by the next item */
    }
}

/* ------------------------------------------------ the three plain kernels */

/* grid (nbatch, 2, nslab) */
template <bool BIG>
__global__ void __launch_bounds__(256, 2) k_chain_B(NsbGeom g, ChainArgs a) {
    NSB_DYN_SMEM(cd, tile);
    chain_B_item<BIG>(g, a, tile, a.slab0 + blockIdx.z, (a.slab0 + blockIdx.z) % a.nslotZ, blockIdx.y, blockIdx.x);
}

/* grid (groups per slab, nslab), W warps */
template <bool UMAX>
__global__ void __launch_bounds__(256, 1) k_chain_G(NsbGeom g, ChainArgs a) {
    NSB_DYN_SMEM(unsigned char, raw);
    cd *TW = reinterpret_cast<cd *>(raw);
    nsb_fill_tw1024(TW, g.TWX);
    __syncthreads();
    const int slab = a.slab0 + blockIdx.y;
    chain_G_item<UMAX>(a, raw + NSB_CH_TW_BYTES, TW, slab % a.nslotZ, slab % a.nslotP, blockIdx.x);
}

/* persistent G: grid = SMs, every warp walks pairs warp_global + k * total_warps of all slabs, so the
 * twiddle table is built once per SM and the first (otherwise cold) input of the next pair is pulled
 * towards L2 while the forward transform of the current one runs */
template <bool UMAX>
__global__ void __launch_bounds__(256, 1) k_chain_Gp(NsbGeom g, ChainArgs a) {
    NSB_DYN_SMEM(unsigned char, raw);
    cd *TW = reinterpret_cast<cd *>(raw);
    nsb_fill_tw1024(TW, g.TWX);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int NP = a.NQ >> 1;
    const long long total = (long long)a.nslab * NP, stride = (long long)gridDim.x * W;
    unsigned char *wb = raw + NSB_CH_TW_BYTES + (size_t)warp * NSB_CHG_WARP_BYTES(UMAX);
    double umax = 0.0;
    for (long long p = (long long)blockIdx.x * W + warp; p < total; p += stride) {
        const int sl = (int)(p / NP), pair = (int)(p - (long long)sl * NP);
        const int slab = a.slab0 + sl;
        const cd *z1 = a.Z + ((size_t)((slab % a.nslotZ) * 2 + 0) * a.NQ + 2 * pair) * 1024;
        const cd *z2 = a.Z + ((size_t)((slab % a.nslotZ) * 2 + 1) * a.NQ + 2 * pair) * 1024;
        cd *po = a.P + ((size_t)(slab % a.nslotP) * NP + pair) * 1024;
        const cd *zn = nullptr;
        if (p + stride < total) {
            const long long q = p + stride;
            const int s2 = a.slab0 + (int)(q / NP), p2 = (int)(q - (long long)(q / NP) * NP);
            zn = a.Z + ((size_t)((s2 % a.nslotZ) * 2 + 0) * a.NQ + 2 * p2) * 1024;
        }
        const double u = chain_G_pair<UMAX>(z1, z2, po, wb, TW, lane, zn);
        umax = nsb_max_sticky(umax, u);
    }
    if (UMAX) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) umax = nsb_max_sticky(umax, __shfl_xor_sync(0xffffffffu, umax, off));
        if (lane == 0) atomicMax(a.umax_bits, (unsigned long long)__double_as_longlong(umax));
    }
}


/*
 * G with tensor-memory parking and TMA-fed inputs (the default on hardware). Same arithmetic as chain_G_pair
 * (bit-identical results), different data movement:
 *   - persistent, one CTA per SM, 8 independent warps (no block barrier after start-up), also with UMAX
 *     (the shared-memory version had room for 6 warps there);
 *   - every 16 KB comb spectrum is fetched by ONE cp.async.bulk into the warp's landing buffer, signalled on the
 *     warp's mbarrier; the fetch of input n+1 is issued as soon as input n sits in registers, so it flies during
 *     the whole transform of input n (round 1: 35-40 % of G's stall samples were global-load waits / LSU throttle
 *     behind 32 back-to-back LDG.128 per lane);
 *   - the per-lane values that must survive a register-resident transform (Re*Im products, the Jacobian of comb 0,
 *     psi_x for max|u|) are parked in TENSOR MEMORY (tcgen05.st / tcgen05.ld, 192 of the warp's 256 columns), not
 *     in 16-24 KB of shared memory per warp -- which is what makes room for the landing buffers.
 */
#define NSB_CHG2_WARP_BYTES (16384 + 8192)
#define NSB_CHG2_SMEM (NSB_CH_TW_BYTES + 8 * NSB_CHG2_WARP_BYTES + 8 * 8 + 16)

template <bool UMAX>
__global__ void __launch_bounds__(256, 1) k_chain_G2(NsbGeom g, ChainArgs a) {
    NSB_DYN_SMEM(unsigned char, raw);
    cd *TW = reinterpret_cast<cd *>(raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wb = raw + NSB_CH_TW_BYTES + (size_t)warp * NSB_CHG2_WARP_BYTES;
    cd *L = reinterpret_cast<cd *>(wb);
    double *T = reinterpret_cast<double *>(wb + 16384);
    nsb_mbar_t *bars = reinterpret_cast<nsb_mbar_t *>(raw + NSB_CH_TW_BYTES + 8 * NSB_CHG2_WARP_BYTES);
    unsigned *tslot = reinterpret_cast<unsigned *>(bars + 8);
    nsb_mbar_t *bar = &bars[warp];
    nsb_fill_tw1024(TW, g.TWX);
    if (lane == 0) nsb_mbar_init(bar, 1);
    const unsigned tbase = nsb_tmem_alloc_all(tslot); /* block barrier inside: TW, mbarriers, TMEM ready */
    /* this warp's 256 columns: [0,64) J of comb 0 | [64,128) Re*Im of the current comb | [128,192) psi_x */
    const unsigned tcol = nsb_tmem_warp_base(tbase, warp, (warp >> 2) * 256);
    /* pairs per slab NP = NQ / 2 is a power of two; one launch covers at most 2^17 pairs; the plain chain keeps
     * every slab in its own Z / P slot (nslotZ == nslotP == number of slabs of the device) */
    const int NP = a.NQ >> 1, nps = nsb_log2i(NP);
    const int total = a.nslab * NP, stride = (int)gridDim.x * 8;
    auto zsrc = [&](int q, int comb, int pp) -> const cd * {
        const int slab = a.slab0 + (q >> nps), pair = q & (NP - 1);
        return a.Z + ((size_t)(slab * 2 + pp) * a.NQ + 2 * pair + comb) * 1024;
    };
    auto issue = [&](const cd *src) {
        if (lane == 0) {
            nsb_fence_proxy_async();
            nsb_mbar_expect_tx(bar, 16384u);
            nsb_bulk_g2s(L, src, 16384u, bar);
        }
    };
    unsigned phase = 0u;
    int p = (int)blockIdx.x * 8 + warp;
    if (p < total) issue(zsrc(p, 0, 0));
    double umax = 0.0;
    for (; p < total; p += stride) {
        cd x[32];
#pragma unroll 1
        for (int comb = 0; comb < 2; ++comb) {
            nsb_mbar_wait(bar, phase);
            phase ^= 1u;
#pragma unroll
            for (int m = 0; m < 32; ++m) x[m] = L[lane + 32 * m];
            __syncwarp(); /* every lane has its values: the landing buffer may be refilled */
            issue(zsrc(p, comb, 1));
            nsb_fft1024_warp_sw<true>(x, T, TW, lane); /* x = psi_x + i omega_y */
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                double t[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) t[i] = __dmul_rn(x[8 * h + i].x, x[8 * h + i].y);
                nsb_tmem_st8(tcol + 64 + 16 * h, t);
                if (UMAX) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) t[i] = x[8 * h + i].x;
                    nsb_tmem_st8(tcol + 128 + 16 * h, t);
                }
            }
            nsb_tmem_wait_st();
            nsb_mbar_wait(bar, phase);
            phase ^= 1u;
#pragma unroll
            for (int m = 0; m < 32; ++m) x[m] = L[lane + 32 * m];
            __syncwarp();
            if (comb == 0) issue(zsrc(p, 1, 0));
            else if (p + stride < total) issue(zsrc(p + stride, 0, 0));
            nsb_fft1024_warp_sw<true>(x, T, TW, lane); /* x = psi_y + i omega_x */
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                double t[8], j[8];
                nsb_tmem_ld8(t, tcol + 64 + 16 * h);
#pragma unroll
                for (int i = 0; i < 8; ++i) /* spectral_core.c:103 : J = psi_x*omega_y - psi_y*omega_x */
                    j[i] = __dsub_rn(t[i], __dmul_rn(x[8 * h + i].x, x[8 * h + i].y));
                if (UMAX) { /* spectral_time_integration.c:16 : |u|^2 = psi_y^2 + psi_x^2 */
                    nsb_tmem_ld8(t, tcol + 128 + 16 * h);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        umax = nsb_max_sticky(umax, __dadd_rn(__dmul_rn(x[8 * h + i].x, x[8 * h + i].x), __dmul_rn(t[i], t[i])));
                }
                if (comb == 0) {
                    nsb_tmem_st8(tcol + 16 * h, j);
                } else {
                    nsb_tmem_ld8(t, tcol + 16 * h);
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[8 * h + i] = cmk(t[i], j[i]);
                }
            }
            nsb_tmem_wait_st();
        }
        nsb_fft1024_warp_sw<false>(x, T, TW, lane);
        cd *po = a.P + ((size_t)(a.slab0 + (p >> nps)) * NP + (p & (NP - 1))) * 1024;
#pragma unroll
        for (int m = 0; m < 32; ++m) po[lane + 32 * m] = x[m];
    }
    if (UMAX) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) umax = nsb_max_sticky(umax, __shfl_xor_sync(0xffffffffu, umax, off));
        if (lane == 0) atomicMax(a.umax_bits, (unsigned long long)__double_as_longlong(umax));
    }
    __syncthreads(); /* every warp is done with tensor memory */
    nsb_tmem_free_all(tbase);
}

/* grid (nbatch, nslab) */
template <bool BIG>
__global__ void __launch_bounds__(256) k_chain_F(NsbGeom g, ChainArgs a) {
    NSB_DYN_SMEM(cd, tile);
    chain_F_item<BIG>(g, a, tile, a.slab0 + blockIdx.y, (a.slab0 + blockIdx.y) % a.nslotP, blockIdx.x);
}

/* ------------------------------------------------ the persistent dataflow kernel */

__device__ __forceinline__ void chain_wait(const unsigned *ctr, unsigned need) {
    if (threadIdx.x == 0) {
        while (nsb_ld_acquire(ctr) < need) nsb_backoff();
    }
    __syncthreads();
}

__device__ __forceinline__ void chain_signal(unsigned *ctr) {
    __syncthreads();
    if (threadIdx.x == 0) nsb_red_release(ctr, 1u);
}

template <bool UMAX, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_chain(NsbGeom g, ChainArgs a) {
    NSB_DYN_SMEM(unsigned char, raw);
    __shared__ unsigned s_ticket[2];
    cd *TW = reinterpret_cast<cd *>(raw);
    unsigned char *ub = raw + NSB_CH_TW_BYTES;
    cd *tile = reinterpret_cast<cd *>(ub);
    nsb_fill_tw1024(TW, g.TWX);
    const int W = blockDim.x >> 5;
    const int NP = a.NQ >> 1;
    const int nB = 2 * a.nbatch, nG = (NP + W - 1) / W, nF = a.nbatch;
    const int per = nF + nG + nB;
    const unsigned total = (unsigned)(a.nslab + a.leadG + a.leadF) * (unsigned)per;
    unsigned *doneB = a.ctr + 1, *doneG = a.ctr + 1 + a.nslab, *doneF = a.ctr + 1 + 2 * a.nslab;
    if (threadIdx.x == 0) s_ticket[0] = atomicAdd(a.ctr, 1u);
    __syncthreads();
    for (int it = 0;; ++it) {
        const unsigned ticket = s_ticket[it & 1];
        if (ticket >= total) break;
        /* the next ticket is drawn now so that its latency hides behind this item */
        if (threadIdx.x == 0) s_ticket[(it + 1) & 1] = atomicAdd(a.ctr, 1u);
        const int step = (int)(ticket / (unsigned)per);
        int idx = (int)(ticket - (unsigned)step * (unsigned)per);
        if (idx < nF) { /* F of slab step - leadG - leadF */
            const int s = step - a.leadG - a.leadF;
            if (s >= 0 && s < a.nslab) {
                chain_wait(&doneG[s], (unsigned)nG);
                chain_F_item<false>(g, a, tile, s, s % a.nslotP, idx);
                chain_signal(&doneF[s]);
            }
        } else if (idx < nF + nG) { /* G of slab step - leadG */
            const int s = step - a.leadG;
            if (s >= 0 && s < a.nslab) {
                chain_wait(&doneB[s], (unsigned)nB);
                if (s >= a.nslotP) chain_wait(&doneF[s - a.nslotP], (unsigned)nF);
                chain_G_item<UMAX>(a, ub, TW, s % a.nslotZ, s % a.nslotP, idx - nF);
                chain_signal(&doneG[s]);
            }
        } else { /* B of slab step */
            const int s = step;
            if (s < a.nslab) {
                if (s >= a.nslotZ) chain_wait(&doneG[s - a.nslotZ], (unsigned)nG);
                idx -= nF + nG;
                chain_B_item<false>(g, a, tile, s, s % a.nslotZ, idx / a.nbatch, idx % a.nbatch);
                chain_signal(&doneB[s]);
            }
        }
        __syncthreads(); /* tile / per-warp buffers are reused by the next item; s_ticket visible */
    }
}

#endif
