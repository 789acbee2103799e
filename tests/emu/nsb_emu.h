/*
 * tests/emu/nsb_emu.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A minimal CUDA-on-CPU execution layer so the *same kernel source* that nvcc
 * compiles for sm_100a (navier-stokes-solvers_b200/csrc/cuda/*.cu) can be compiled by
 * g++ (-DNSB_EMULATE) and run in this GPU-less container: every CTA is executed as
 * `blockDim` cooperative fibers (ucontext), __syncthreads() yields to the next fiber,
 * shared memory is one heap block per CTA. It exists to check index algebra, FFT
 * decomposition and barrier placement of the kernels against the CPU oracle in
 * `pytest -m "not gpu"`. It is NOT a product code path: the product loader
 * (navier-stokes-solvers_b200/__init__.py) only ever opens libnsb.so built by nvcc and
 * fails loudly without a GPU; libnsb_emu.so lives under tests/emu/build/ and is opened
 * by tests only. Set NSB_EMU_REVERSE=1 to run fibers in descending thread order
 * (catches some missing-barrier bugs).
 */
#ifndef NSB_EMU_H
#define NSB_EMU_H

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <ucontext.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static /* blocks run one at a time on one OS thread */

struct uint3_emu { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct double2 { double x, y; };
static inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }

namespace nsbemu {
struct Fiber {
    ucontext_t ctx;
    char *stack;
    uint3_emu tid;
    bool done;
    int warp, lane;
};
struct State {
    uint3_emu bid;
    dim3 bdim, gdim;
    unsigned char *smem;
    Fiber *cur;
    ucontext_t sched;
    std::function<void()> *body;
    /* barrier bookkeeping: arrivals are counted, exited threads drop out */
    int live, block_arrived;
    unsigned block_gen;
    std::vector<int> warp_live, warp_arrived;
    std::vector<unsigned> warp_gen;
    std::vector<double> warp_buf;
    int named_arrived[16];
    unsigned named_gen[16];
    uint32_t tmem[128][512]; /* tensor memory of the current CTA (nsb_platform.h: nsb_tmem_*) */
};
extern State g;
void launch(dim3 grid, dim3 block, size_t smem_bytes, std::function<void()> body);
void barrier();
void named_barrier(int id, int nthreads);
void warp_barrier();
void yield();
double shfl(double v, int src_lane);
}  // namespace nsbemu

#define threadIdx (nsbemu::g.cur->tid)
#define blockIdx (nsbemu::g.bid)
#define blockDim (nsbemu::g.bdim)
#define gridDim (nsbemu::g.gdim)
#define __syncthreads() nsbemu::barrier()
#define __syncwarp() nsbemu::warp_barrier()
static inline double __shfl_xor_sync(unsigned, double v, int lane_mask) {
    return nsbemu::shfl(v, nsbemu::g.cur->lane ^ lane_mask);
}
static inline double __shfl_sync(unsigned, double v, int src_lane) { return nsbemu::shfl(v, src_lane); }
#define NSB_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(nsbemu::g.smem)

template <class T> static inline T __ldg(const T *p) { return *p; }
static inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
static inline double __longlong_as_double(long long l) { double r; memcpy(&r, &l, 8); return r; }
static inline unsigned long long atomicMax(unsigned long long *a, unsigned long long v) {
    unsigned long long o = *a; if (v > o) *a = v; return o;
}
static inline double atomicAdd(double *a, double v) { double o = *a; *a = o + v; return o; }
static inline int atomicAdd(int *a, int v) { int o = *a; *a = o + v; return o; }
static inline unsigned atomicAdd(unsigned *a, unsigned v) { unsigned o = *a; *a = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *a, unsigned long long v) { unsigned long long o = *a; *a = o + v; return o; }
static inline unsigned long long atomicXor(unsigned long long *a, unsigned long long v) { unsigned long long o = *a; *a = o ^ v; return o; }
static inline unsigned atomicOr(unsigned *a, unsigned v) { unsigned o = *a; *a = o | v; return o; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }

/* ---- tiny CUDA runtime stand-in (host side of the context code) ---- */
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaSharedmemCarveoutMaxShared = 100 };
enum { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = 2; return 0; }
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
static inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h,
                                            cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; ++r) memmove((char *)d + r * dp, (const char *)s + r * sp, w);
    return 0;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
enum { cudaErrorPeerAccessAlreadyEnabled = 704 };
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return 0; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return 0; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }

#define NSB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    nsbemu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })

#endif
