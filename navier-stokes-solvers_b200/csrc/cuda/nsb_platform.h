/*
 * nsb_platform.h -- one switch between the real target (nvcc, sm_100a) and the
 * CPU kernel-logic emulator used ONLY by tests (tests/emu/nsb_emu.h, -DNSB_EMULATE).
 * The product library libnsb.so is always built by nvcc; nothing in the product
 * selects the emulator at run time.
 */
#ifndef NSB_PLATFORM_H
#define NSB_PLATFORM_H

#ifdef NSB_EMULATE
#include "nsb_emu.h"
#define NSB_HD
#else
#include <cuda_runtime.h>
#define NSB_HD __host__ __device__
#define NSB_DYN_SMEM(type, name)                                   \
    extern __shared__ __align__(16) unsigned char nsb_smem_raw[];  \
    type *name = reinterpret_cast<type *>(nsb_smem_raw)
#define NSB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

typedef double2 cd; /* complex double: x = re, y = im */

/*
 * Inter-CTA dataflow helpers of the persistent chain kernel (nsb_chain.cuh): L2-only loads
 * for data another CTA of the same launch produced (L1 is not coherent), acquire polling and
 * release signalling of completion counters in global memory.
 */
#ifdef NSB_EMULATE
static inline cd nsb_ldcg(const cd *p) { return *p; }
static inline cd nsb_ldcs(const cd *p) { return *p; }
static inline void nsb_stcs(cd *p, cd v) { *p = v; }
static inline void nsb_prefetch_l2(const void *) {}
static inline unsigned nsb_ld_acquire(const unsigned *p) { return *p; }
static inline void nsb_red_release(unsigned *p, unsigned v) { *p += v; }
static inline void nsb_backoff() { nsbemu::yield(); }
static inline void nsb_fence_system() {}
#else
__device__ __forceinline__ cd nsb_ldcg(const cd *p) { return __ldcg(p); }
/* read-once / write-once streams (evict-first) so they do not displace the L2-resident rings */
__device__ __forceinline__ cd nsb_ldcs(const cd *p) { return __ldcs(p); }
__device__ __forceinline__ void nsb_stcs(cd *p, cd v) { __stcs(p, v); }
__device__ __forceinline__ void nsb_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned nsb_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void nsb_red_release(unsigned *p, unsigned v) {
    __threadfence();
    atomicAdd(p, v);
}
__device__ __forceinline__ void nsb_backoff() { __nanosleep(64); }
__device__ __forceinline__ void nsb_fence_system() { __threadfence_system(); }
#endif

/* n / d for CTA-uniform runtime divisors without the I2F/MUFU division sequence (10 % of the stall samples of
 * the round-1 B kernel sat on it): m = ceil(2^32 / d), exact while n * d < 2^32. d == 1 is encoded as m == 0. */
struct NsbFastDiv {
    unsigned d, m;
};
NSB_HD inline NsbFastDiv nsb_fastdiv_make(unsigned d) {
    NsbFastDiv f;
    f.d = d;
    f.m = d <= 1 ? 0u : (unsigned)((0x100000000ull + d - 1) / d);
    return f;
}
#ifdef NSB_EMULATE
static inline unsigned nsb_fastdiv(unsigned n, NsbFastDiv f) { return f.m ? (unsigned)(((unsigned long long)n * f.m) >> 32) : n; }
static inline int nsb_log2i(int v) { return 31 - __builtin_clz((unsigned)v); }
#else
__device__ __forceinline__ unsigned nsb_fastdiv(unsigned n, NsbFastDiv f) { return f.m ? __umulhi(n, f.m) : n; }
__device__ __forceinline__ int nsb_log2i(int v) { return 31 - __clz(v); }
#endif

/* barrier among the `nthreads` consecutive threads of group `id - 1` of the CTA (bar.sync id, nthreads): the two
 * independent 4-warp halves of the two-group column kernel synchronise separately */
#ifdef NSB_EMULATE
static inline void nsb_group_barrier(int id, int nthreads) { nsbemu::named_barrier(id, nthreads); }
#else
__device__ __forceinline__ void nsb_group_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif

/* correctly rounded FP64 reciprocal (MUFU.RCP64H + Newton, inlined: no division subroutine) */
#ifdef NSB_EMULATE
static inline double nsb_drcp(double a) { return 1.0 / a; }
#else
__device__ __forceinline__ double nsb_drcp(double a) { return __drcp_rn(a); }
#endif

/*
 * TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers. global -> shared,
 * completion counted in bytes on an mbarrier in shared memory. Addresses and sizes must be
 * multiples of 16 bytes. Under the CPU emulator the copy is an immediate memcpy.
 */
#ifdef NSB_EMULATE
typedef unsigned long long nsb_mbar_t;
/* emulated mbarrier: low 32 bits = completed phases, high 32 bits = pending tx bytes */
static inline void nsb_mbar_init(nsb_mbar_t *bar, int) { *bar = 0; }
static inline void nsb_mbar_expect_tx(nsb_mbar_t *bar, unsigned bytes) {
    *bar += (unsigned long long)bytes << 32;
    if (bytes == 0) *bar += 1; /* arrive.expect_tx with nothing pending completes the phase */
}
static inline void nsb_bulk_g2s(void *dst, const void *src, unsigned bytes, nsb_mbar_t *bar) {
    memcpy(dst, src, bytes);
    *bar -= (unsigned long long)bytes << 32;
    if ((*bar >> 32) == 0) *bar += 1;
}
static inline void nsb_mbar_wait(nsb_mbar_t *bar, unsigned parity) {
    while (((unsigned)(*bar) & 1u) == parity) nsbemu::yield();
}
static inline void nsb_fence_proxy_async() {}
static inline unsigned nsb_smem_align_pad(const void *p, unsigned a) { return (unsigned)((a - ((size_t)p & (a - 1))) & (a - 1)); }
#else
typedef unsigned long long nsb_mbar_t;
__device__ __forceinline__ unsigned nsb_smem_addr(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void nsb_mbar_init(nsb_mbar_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nsb_smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void nsb_mbar_expect_tx(nsb_mbar_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nsb_smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void nsb_bulk_g2s(void *dst, const void *src, unsigned bytes, nsb_mbar_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     nsb_smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(nsb_smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void nsb_mbar_wait(nsb_mbar_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "NSB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra NSB_DONE_%=;\n"
        "bra NSB_WAIT_%=;\n"
        "NSB_DONE_%=:\n"
        "}\n" ::"r"(nsb_smem_addr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void nsb_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
/* bytes to add to a shared-memory pointer to reach the next multiple of `a` (a power of two) */
__device__ __forceinline__ unsigned nsb_smem_align_pad(const void *p, unsigned a) {
    return (a - (nsb_smem_addr(p) & (a - 1))) & (a - 1);
}
#endif

/*
 * TMA tiled tensor copies (cp.async.bulk.tensor.2d, SASS UTMALDG): one instruction moves a
 * box of box1 rows x box0 doubles from a pitched 2-D FP64 array into shared memory, rows
 * packed back to back; out-of-range coordinates are zero-filled. The descriptor is encoded
 * on the host (nsb_ctx.cu: make_tmap_2d via cuTensorMapEncodeTiled) and passed to the kernel
 * as a __grid_constant__ parameter. Under the CPU emulator it is a plain struct and the copy
 * an immediate strided memcpy.
 */
#ifdef NSB_EMULATE
struct nsb_tmap {
    const double *base;
    long long dim0, dim1;     /* extent in doubles / rows */
    long long stride1;        /* row pitch in bytes */
    int box0, box1;
    /* 3-D maps (nsb_tma_3d_s2g): extents / byte strides of dims 1, 2, box of dim 2, 128B swizzle */
    long long dim2, stride2;
    int box2, swz;
};
#define NSB_GRID_CONSTANT
static inline void nsb_tma_2d_g2s(void *dst, const nsb_tmap *m, int c0, int c1, nsb_mbar_t *bar) {
    double *d = (double *)dst;
    for (int r = 0; r < m->box1; ++r)
        for (int e = 0; e < m->box0; ++e) {
            const long long row = (long long)c1 + r, col = (long long)c0 + e;
            double v = 0.0;
            if (row >= 0 && row < m->dim1 && col >= 0 && col < m->dim0)
                v = *(const double *)((const char *)m->base + row * m->stride1 + col * 8);
            d[(size_t)r * m->box0 + e] = v;
        }
    const unsigned long long bytes = (unsigned long long)m->box0 * m->box1 * 8ull;
    *bar -= bytes << 32;
    if ((*bar >> 32) == 0) *bar += 1;
}
/* shared -> global store of a {box0, 1, box2} box; smem rows are 128 bytes (box0 == 16 doubles) with
 * the 16-byte chunk index XORed with (128-byte line index mod 8), i.e. TMA SWIZZLE_128B */
static inline void nsb_tma_3d_s2g(const nsb_tmap *m, const void *src, int c0, int c1, int c2) {
    const char *s = (const char *)src;
    for (int r = 0; r < m->box2; ++r)
        for (int e = 0; e < m->box0; ++e) {
            const long long i2 = (long long)c2 + r, i1 = c1, i0 = (long long)c0 + e;
            if (i2 < 0 || i2 >= m->dim2 || i1 < 0 || i1 >= m->dim1 || i0 < 0 || i0 >= m->dim0) continue;
            const size_t line = ((size_t)(s + (size_t)r * m->box0 * 8) >> 7) & (m->swz == 2 ? 3 : 7);
            const int chunk = e >> 1;
            const int pch = m->swz ? (chunk ^ (int)line) : chunk;
            const double v = *(const double *)(s + (size_t)r * m->box0 * 8 + (size_t)pch * 16 + (e & 1) * 8);
            *(double *)((char *)m->base + i2 * m->stride2 + i1 * m->stride1 + i0 * 8) = v;
        }
}
/* global -> shared load of a {box0, 1, box2} box, same swizzle convention */
static inline void nsb_tma_3d_g2s(void *dst, const nsb_tmap *m, int c0, int c1, int c2, nsb_mbar_t *bar) {
    char *d = (char *)dst;
    for (int r = 0; r < m->box2; ++r)
        for (int e = 0; e < m->box0; ++e) {
            const long long i2 = (long long)c2 + r, i1 = c1, i0 = (long long)c0 + e;
            double v = 0.0;
            if (i2 >= 0 && i2 < m->dim2 && i1 >= 0 && i1 < m->dim1 && i0 >= 0 && i0 < m->dim0)
                v = *(const double *)((const char *)m->base + i2 * m->stride2 + i1 * m->stride1 + i0 * 8);
            const size_t line = ((size_t)(d + (size_t)r * m->box0 * 8) >> 7) & (m->swz == 2 ? 3 : 7);
            const int chunk = e >> 1;
            const int pch = m->swz ? (chunk ^ (int)line) : chunk;
            *(double *)(d + (size_t)r * m->box0 * 8 + (size_t)pch * 16 + (e & 1) * 8) = v;
        }
    const unsigned long long bytes = (unsigned long long)m->box0 * m->box2 * 8ull;
    *bar -= bytes << 32;
    if ((*bar >> 32) == 0) *bar += 1;
}
static inline void nsb_tma_store_commit() {}
static inline void nsb_tma_store_wait_read() {}
#else
#include <cuda.h>
typedef CUtensorMap nsb_tmap;
#define NSB_GRID_CONSTANT __grid_constant__
__device__ __forceinline__ void nsb_tma_2d_g2s(void *dst, const nsb_tmap *map, int c0, int c1, nsb_mbar_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            nsb_smem_addr(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(nsb_smem_addr(bar))
        : "memory");
}
__device__ __forceinline__ void nsb_tma_3d_g2s(void *dst, const nsb_tmap *map, int c0, int c1, int c2, nsb_mbar_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            nsb_smem_addr(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(nsb_smem_addr(bar))
        : "memory");
}
/* shared -> global tensor store (SASS UTMASTG), tracked by bulk async-groups: commit after
 * issuing, wait_read before the shared-memory source is overwritten */
__device__ __forceinline__ void nsb_tma_3d_s2g(const nsb_tmap *map, const void *src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(c2), "r"(nsb_smem_addr(src))
                 : "memory");
}
__device__ __forceinline__ void nsb_tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void nsb_tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
#endif


/*
 * Tensor memory (TMEM, 256 KB per SM: 128 lanes x 512 columns x 32 bit) used as a LANE-PRIVATE PARKING LOT:
 * the comb kernel parks the real products / Jacobians it has to keep across a register-resident FFT-1024
 * there (tcgen05.st / tcgen05.ld, SASS STTM / LDTM) instead of in shared memory, which frees 128-192 KB of
 * shared memory per SM for TMA landing buffers. A warp may touch only the 32 lanes of its own quadrant
 * (warp % 4); thread t of the warp owns lane 32 * (warp % 4) + t; columns are 32-bit, one double = 2 columns.
 * No tensor-core instruction is involved. Under the CPU emulator TMEM is a plain per-CTA array.
 */
#ifdef NSB_EMULATE
static inline unsigned nsb_tmem_alloc_all(unsigned *) { /* like the real one: contains a block barrier */
    if (nsbemu::g.cur->warp == 0 && nsbemu::g.cur->lane == 0) memset(nsbemu::g.tmem, 0xff, sizeof(nsbemu::g.tmem));
    nsbemu::barrier();
    return 0u;
}
static inline void nsb_tmem_free_all(unsigned) {}
static inline unsigned nsb_tmem_warp_base(unsigned base, int warp, int col) {
    return base + (((unsigned)(32 * (warp & 3))) << 16) + (unsigned)col;
}
template <int N>
static inline void nsb_tmem_stN(unsigned taddr, const double (&v)[N]) {
    const unsigned lane = (taddr >> 16) + (unsigned)nsbemu::g.cur->lane, col = taddr & 0xffffu;
    memcpy(&nsbemu::g.tmem[lane][col], v, 8 * N);
}
template <int N>
static inline void nsb_tmem_ldN(double (&v)[N], unsigned taddr) {
    const unsigned lane = (taddr >> 16) + (unsigned)nsbemu::g.cur->lane, col = taddr & 0xffffu;
    memcpy(v, &nsbemu::g.tmem[lane][col], 8 * N);
}
static inline void nsb_tmem_st16(unsigned t, const double (&v)[16]) { nsb_tmem_stN<16>(t, v); }
static inline void nsb_tmem_ld16(double (&v)[16], unsigned t) { nsb_tmem_ldN<16>(v, t); }
static inline void nsb_tmem_st8(unsigned t, const double (&v)[8]) { nsb_tmem_stN<8>(t, v); }
static inline void nsb_tmem_ld8(double (&v)[8], unsigned t) { nsb_tmem_ldN<8>(v, t); }
static inline void nsb_tmem_wait_st() {}
static inline void nsb_tmem_wait_ld() {}
struct NsbTmemRaw8 { double d[8]; };
static inline void nsb_tmem_ld8_async(NsbTmemRaw8 &q, unsigned t) { nsb_tmem_ldN<8>(q.d, t); }
static inline void nsb_tmem_wait_ld8(double (&v)[8], NsbTmemRaw8 &q) { memcpy(v, q.d, sizeof(q.d)); }
#else
/* one warp allocates all 512 columns (one CTA per SM by construction), every thread gets the base address */
__device__ __forceinline__ unsigned nsb_tmem_alloc_all(unsigned *slot_smem) {
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(nsb_smem_addr(slot_smem))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *reinterpret_cast<volatile unsigned *>(slot_smem);
}
/* all warps must be done with TMEM (caller synchronises the CTA first) */
__device__ __forceinline__ void nsb_tmem_free_all(unsigned base) {
    if ((threadIdx.x >> 5) == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}
__device__ __forceinline__ unsigned nsb_tmem_warp_base(unsigned base, int warp, int col) {
    return base + (((unsigned)(32 * (warp & 3))) << 16) + (unsigned)col;
}
/* N doubles of this thread <-> 2N consecutive columns of its own lane, starting at column (taddr & 0xffff) */
__device__ __forceinline__ void nsb_tmem_st8(unsigned taddr, const double (&v)[8]) {
    asm volatile(
        "{\n"
        ".reg .b32 a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15;\n"
        "mov.b64 {a0, a1}, %1;\n"
        "mov.b64 {a2, a3}, %2;\n"
        "mov.b64 {a4, a5}, %3;\n"
        "mov.b64 {a6, a7}, %4;\n"
        "mov.b64 {a8, a9}, %5;\n"
        "mov.b64 {a10, a11}, %6;\n"
        "mov.b64 {a12, a13}, %7;\n"
        "mov.b64 {a14, a15}, %8;\n"
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15};\n"
        "}\n" ::"r"(taddr),
        "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "d"(v[4]), "d"(v[5]), "d"(v[6]), "d"(v[7])
        : "memory");
}
__device__ __forceinline__ void nsb_tmem_ld8(double (&v)[8], unsigned taddr) {
    asm volatile(
        "{\n"
        ".reg .b32 a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15;\n"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15}, [%8];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        "mov.b64 %0, {a0, a1};\n"
        "mov.b64 %1, {a2, a3};\n"
        "mov.b64 %2, {a4, a5};\n"
        "mov.b64 %3, {a6, a7};\n"
        "mov.b64 %4, {a8, a9};\n"
        "mov.b64 %5, {a10, a11};\n"
        "mov.b64 %6, {a12, a13};\n"
        "mov.b64 %7, {a14, a15};\n"
        "}\n"
        : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]), "=d"(v[4]), "=d"(v[5]), "=d"(v[6]), "=d"(v[7])
        : "r"(taddr)
        : "memory");
}
/* split form of nsb_tmem_ld8 so that the next load can be in flight while the previous values are used: the raw
 * 32-bit registers must not be read before nsb_tmem_wait_ld8 (which carries them as in/out operands, so no use
 * can move above it) has turned them into doubles */
struct NsbTmemRaw8 { unsigned r[16]; };
__device__ __forceinline__ void nsb_tmem_ld8_async(NsbTmemRaw8 &q, unsigned taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(q.r[0]), "=r"(q.r[1]), "=r"(q.r[2]), "=r"(q.r[3]), "=r"(q.r[4]), "=r"(q.r[5]), "=r"(q.r[6]), "=r"(q.r[7]),
                   "=r"(q.r[8]), "=r"(q.r[9]), "=r"(q.r[10]), "=r"(q.r[11]), "=r"(q.r[12]), "=r"(q.r[13]), "=r"(q.r[14]), "=r"(q.r[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void nsb_tmem_wait_ld8(double (&v)[8], NsbTmemRaw8 &q) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(q.r[0]), "+r"(q.r[1]), "+r"(q.r[2]), "+r"(q.r[3]), "+r"(q.r[4]), "+r"(q.r[5]), "+r"(q.r[6]), "+r"(q.r[7]),
                   "+r"(q.r[8]), "+r"(q.r[9]), "+r"(q.r[10]), "+r"(q.r[11]), "+r"(q.r[12]), "+r"(q.r[13]), "+r"(q.r[14]), "+r"(q.r[15])
                 :
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __hiloint2double((int)q.r[2 * i + 1], (int)q.r[2 * i]);
}
__device__ __forceinline__ void nsb_tmem_st16(unsigned taddr, const double (&v)[16]) {
    asm volatile(
        "{\n"
        ".reg .b32 a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15, a16, a17, a18, a19, a20, a21, a22, a23, a24, a25, a26, a27, a28, a29, a30, a31;\n"
        "mov.b64 {a0, a1}, %1;\n"
        "mov.b64 {a2, a3}, %2;\n"
        "mov.b64 {a4, a5}, %3;\n"
        "mov.b64 {a6, a7}, %4;\n"
        "mov.b64 {a8, a9}, %5;\n"
        "mov.b64 {a10, a11}, %6;\n"
        "mov.b64 {a12, a13}, %7;\n"
        "mov.b64 {a14, a15}, %8;\n"
        "mov.b64 {a16, a17}, %9;\n"
        "mov.b64 {a18, a19}, %10;\n"
        "mov.b64 {a20, a21}, %11;\n"
        "mov.b64 {a22, a23}, %12;\n"
        "mov.b64 {a24, a25}, %13;\n"
        "mov.b64 {a26, a27}, %14;\n"
        "mov.b64 {a28, a29}, %15;\n"
        "mov.b64 {a30, a31}, %16;\n"
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15, a16, a17, a18, a19, a20, a21, a22, a23, a24, a25, a26, a27, a28, a29, a30, a31};\n"
        "}\n" ::"r"(taddr),
        "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "d"(v[4]), "d"(v[5]), "d"(v[6]), "d"(v[7]), "d"(v[8]), "d"(v[9]), "d"(v[10]), "d"(v[11]), "d"(v[12]), "d"(v[13]), "d"(v[14]), "d"(v[15])
        : "memory");
}
__device__ __forceinline__ void nsb_tmem_ld16(double (&v)[16], unsigned taddr) {
    asm volatile(
        "{\n"
        ".reg .b32 a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15, a16, a17, a18, a19, a20, a21, a22, a23, a24, a25, a26, a27, a28, a29, a30, a31;\n"
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15, a16, a17, a18, a19, a20, a21, a22, a23, a24, a25, a26, a27, a28, a29, a30, a31}, [%16];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        "mov.b64 %0, {a0, a1};\n"
        "mov.b64 %1, {a2, a3};\n"
        "mov.b64 %2, {a4, a5};\n"
        "mov.b64 %3, {a6, a7};\n"
        "mov.b64 %4, {a8, a9};\n"
        "mov.b64 %5, {a10, a11};\n"
        "mov.b64 %6, {a12, a13};\n"
        "mov.b64 %7, {a14, a15};\n"
        "mov.b64 %8, {a16, a17};\n"
        "mov.b64 %9, {a18, a19};\n"
        "mov.b64 %10, {a20, a21};\n"
        "mov.b64 %11, {a22, a23};\n"
        "mov.b64 %12, {a24, a25};\n"
        "mov.b64 %13, {a26, a27};\n"
        "mov.b64 %14, {a28, a29};\n"
        "mov.b64 %15, {a30, a31};\n"
        "}\n"
        : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]), "=d"(v[4]), "=d"(v[5]), "=d"(v[6]), "=d"(v[7]), "=d"(v[8]), "=d"(v[9]), "=d"(v[10]), "=d"(v[11]), "=d"(v[12]), "=d"(v[13]), "=d"(v[14]), "=d"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void nsb_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void nsb_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
#endif

#endif
