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
            const size_t line = ((size_t)(s + (size_t)r * m->box0 * 8) >> 7) & 7;
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
            const size_t line = ((size_t)(d + (size_t)r * m->box0 * 8) >> 7) & 7;
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

#endif
