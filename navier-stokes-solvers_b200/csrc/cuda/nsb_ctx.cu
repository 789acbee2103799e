/*
 * nsb_ctx.cu -- implementation of include/nsb.h: device context, buffers, kernel
 * launches. Host-side C++ kept deliberately thin; everything numerical is in
 * nsb_kernels.cuh / nsb_fft.cuh.
 */
#include "nsb_kernels.cuh"
#include "nsb_chain.cuh"
#include "nsb_stats.cuh"
#include "../../../include/nsb.h"

#ifndef NSB_EMULATE
#include <nvtx3/nvToolsExt.h> /* header-only NVTX 3: ranges show up under nsys / ncu --nvtx, no link dependency */
struct NsbRange {
    explicit NsbRange(const char *name) { nvtxRangePushA(name); }
    ~NsbRange() { nvtxRangePop(); }
};
#else
struct NsbRange {
    explicit NsbRange(const char *) {}
};
#endif

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#define NSB_MIN_DT 1e-8 /* /root/reference/include/common_types.h:32 */
#define NSB_MAX_DT 0.1  /* /root/reference/include/common_types.h:33 */
#define NSB_FORCE_RING 16

static thread_local char g_err[512] = "";

static int fail(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return -1;
}

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

#define CKL()                                                                              \
    do {                                                                                   \
        cudaError_t e_ = cudaGetLastError();                                               \
        if (e_ != cudaSuccess)                                                             \
            return fail("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

struct nsb_ctx {
    nsb_params prm;
    NsbGeom g;
    size_t N, M;
    int device;
    cudaStream_t stream;
    void *fields[NSB_NUM_FIELDS];
    cd *bufA, *bufB;
    size_t buf_elems;
    double *d_kx, *d_ky, *d_kyK;
    cd *d_WY, *d_WX, *d_TW, *d_TWK;
    NsbScalars *d_sc;
    NsbScalars *h_sc; /* pinned */
    NsbHv hv;
    /* forcing */
    cd *d_force;
    size_t force_cap;
    int ib, jb, force_on;
    cd *h_force[NSB_FORCE_RING];
    size_t h_force_cap[NSB_FORCE_RING];
    cudaEvent_t force_ev[NSB_FORCE_RING];
    int force_slot;
    int masked_nonzero;
    int has_beta;
    int fast; /* warp-FFT fast paths enabled (NSB_DISABLE_FAST=1 forces the generic kernels) */
    /* K-layout copies of the RK state used by the fused column kernel (k_alpha1024) */
    cd *kW0, *kWst, *kACC, *kSFD;
    int k_truth; /* 1: kW0 (and kSFD) hold the current state, the reference-layout fields are stale */
    /* chain (nsb_chain.cuh): 0 off, 1 three plain kernels, 2 persistent dataflow kernel */
    int chain_mode;
    int ch_rows; /* y rows per chain slab (ChainArgs::R) */
    int ch_fbc;  /* F batch width override (0 = automatic) */
    int ch_fthreads;
    cd *chZ, *chP;     /* L2-resident rings of the persistent kernel */
    unsigned *chCtr;   /* ticket + per-slab completion counters */
    int ch_slotZ, ch_slotP, ch_leadG, ch_leadF, ch_ctas, ch_btma;
    int alpha_cols; /* columns per CTA of the fused column kernel: 4 (two CTAs per SM) or 8 */
    int alpha_groups;  /* TMA variant: 1 = eight columns per CTA, 2 = two independent 4-column groups per CTA */
    int alpha_persist; /* TMA variant: one CTA per SM walking the tiles (NSB_ALPHA_PERSIST=0: one tile per CTA) */
    int alpha_tma;  /* tile stores of the column kernel through TMA (8 columns per CTA, 128B swizzle) */
    NsbTmaps tmA;   /* tensor maps of every device's T for those stores */
    int tmA_ok;
    nsb_tmap tmT, tmT8; /* tensor maps of T in bufA (TMA-fed B), 16- and 8-column boxes */
    int tmT_ok, tmT8_ok;
    int ch_bbc;
    int ch_gpersist;
    int ch_ftma; /* F persistent and TMA-fed (k_chain_F_tma) */
    nsb_tmap tmP;
    const void *tmP_base; /* the P buffer tmP describes */
    int ch_b512; /* B with split step 2 and twice the threads (k_chain_B_tma<512, 1, false, true>) */
    int ch_g2; /* G with TMEM parking + TMA-fed inputs (k_chain_G2) */
    int ch_pieces;          /* slab ranges of the chain run on two streams (1 = one range, one stream) */
    cudaStream_t stream2;
    cd *bufP;               /* P of the split chain */
    cudaEvent_t ch_ev[4];
    /* slab decomposition over the devices of one NVSwitch domain */
    int mg_rank, mg_world;
    int mg_cb, mg_c0, mg_c1;     /* columns per device, own column range */
    int mg_rows, mg_row_shift;   /* own y rows = ny / world */
    cd *bufTp;                   /* T' exchange buffer [ny][cb] (peers push into it) */
    cd *mg_T[NSB_MAXP], *mg_Tp[NSB_MAXP];
    NsbScalars *mg_sc[NSB_MAXP];
    int mg_ipc;                  /* peers mapped through CUDA IPC (one process per device) */
    unsigned mg_epoch;
    void *mg_opened[3 * NSB_MAXP];
    int mg_nopened;
    cudaEvent_t mg_ev;
    int sm_count;
    long long launches;
    /* CUDA-graph replay of the device-dt step on launch-bound (small) grids */
    int cfl_done;              /* nsb_rk4_step_begin has already launched this step's CFL kernel */
    int graph_mode;            /* NSB_GRAPH: -1 automatic (grids up to 4096^2), 0 off, 1 always */
    int graph_warm;            /* consecutive plain steps taken with the current key */
    long long graph_launches;  /* kernels inside the captured step */
    unsigned char graph_key[192];
#ifndef NSB_EMULATE
    cudaGraphExec_t graph_exec;
#endif
    std::map<const void *, size_t> *smem_set; /* kernel -> dynamic shared memory size its attribute allows */
    cudaEvent_t ev0, ev1;
    /* optional per-kernel CUDA-event timing (nsb_profile_begin/_end) */
    int prof_on;
    std::vector<cudaEvent_t> *prof_ev; /* pairs */
    std::vector<int> *prof_kind;
};

enum { PK_ALPHA_INV = 0, PK_BETA_INV, PK_GAMMA, PK_BETA_FWD, PK_ALPHA_FWD, PK_OTHER, PK_COUNT };

static void prof_pre(nsb_ctx *c, int kind) {
    if (!c->prof_on) return;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    c->prof_ev->push_back(e0);
    c->prof_ev->push_back(e1);
    c->prof_kind->push_back(kind);
    cudaEventRecord(e0, c->stream);
}
static void prof_post(nsb_ctx *c) {
    if (!c->prof_on) return;
    cudaEventRecord(c->prof_ev->back(), c->stream);
}

extern "C" const char *nsb_last_error(void) { return g_err; }

extern "C" int nsb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

static bool pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

static int field_is_phys(int f) { return f >= NSB_U && f < NSB_NUM_FIELDS; }
static int field_is_hat(int f) { return f >= 0 && f < NSB_NUM_SPECTRAL; }

static int get_field(nsb_ctx *c, int f, void **out) {
    if (!field_is_phys(f) && !field_is_hat(f)) return fail("bad field id %d", f);
    if (!c->fields[f]) {
        const size_t bytes = field_is_phys(f) ? c->N * sizeof(double) : c->M * sizeof(cd);
        CK(cudaMalloc(&c->fields[f], bytes));
        CK(cudaMemsetAsync(c->fields[f], 0, bytes, c->stream));
    }
    *out = c->fields[f];
    return 0;
}

template <class K>
static int ensure_smem(nsb_ctx *c, K kernel, size_t bytes, bool max_carveout = false) {
    if (bytes <= 48 * 1024 && !max_carveout) return 0;
    /* the attribute belongs to the function (per device), not to a context: one process-wide table,
     * only ever raised, so contexts with different tile sizes do not undercut each other */
    static std::map<std::pair<int, const void *>, size_t> table;
    static std::mutex table_mutex;
    std::lock_guard<std::mutex> lock(table_mutex);
    const std::pair<int, const void *> key(c->device, (const void *)kernel);
    auto it = table.find(key);
    if (it != table.end() && it->second >= bytes) return 0; /* the attribute already covers this launch */
    if (bytes > 227 * 1024) return fail("kernel needs %zu bytes of shared memory (> 227 KB)", bytes);
    if (bytes > 48 * 1024)
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    /* measured (round 1): forcing the large carve-out on the L1-reliant kernels (alpha, beta)
     * costs 5-15 %, so only kernels that ask for it get it */
    if (max_carveout)
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    table[key] = bytes;
    return 0;
}

static void fill_twiddles(std::vector<cd> &w, int n) {
    const long double two_pi = 6.283185307179586476925286766559005768L;
    w.resize(n);
    for (int k = 0; k < n; ++k) {
        /* octant reduction keeps the long-double argument small */
        const int quad = (int)((4LL * k) / n);
        const int rem = k - (int)(((long long)quad * n) / 4);
        const long double a = two_pi * (long double)rem / (long double)n;
        const long double cs = cosl(a), sn = sinl(a);
        long double re, im;
        switch (quad) {
            case 0: re = cs; im = -sn; break;
            case 1: re = -sn; im = -cs; break;
            case 2: re = -cs; im = sn; break;
            default: re = sn; im = cs; break;
        }
        w[k] = make_double2((double)re, (double)im);
    }
}

extern "C" int nsb_ctx_create(const nsb_params *p, nsb_ctx **out) {
    if (!p || !out) return fail("nsb_ctx_create: null argument");
    *out = nullptr;
    if (!pow2(p->nx) || !pow2(p->ny) || p->nx < 8 || p->ny < 8)
        return fail("nx, ny must be powers of two >= 8 (got %d x %d)", p->nx, p->ny);
    if (!(p->Lx > 0.0) || !(p->Ly > 0.0) || !(p->Re > 0.0)) return fail("Lx, Ly, Re must be positive");
#ifndef NSB_EMULATE
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail("no CUDA device visible: this library has no CPU fallback");
    if (p->device < 0 || p->device >= ndev) return fail("device %d out of range (%d visible)", p->device, ndev);
#endif
    CK(cudaSetDevice(p->device));
    nsb_ctx *c = new nsb_ctx();
    memset(c->fields, 0, sizeof(c->fields));
    c->prm = *p;
    c->device = p->device;
    c->smem_set = new std::map<const void *, size_t>();
    c->launches = 0;
    c->graph_mode = getenv("NSB_GRAPH") ? atoi(getenv("NSB_GRAPH")) : -1;
    c->cfl_done = 0;
    c->graph_warm = 0;
    c->graph_launches = 0;
    memset(c->graph_key, 0, sizeof(c->graph_key));
#ifndef NSB_EMULATE
    c->graph_exec = nullptr;
#endif
    c->prof_on = 0;
    c->prof_ev = new std::vector<cudaEvent_t>();
    c->prof_kind = new std::vector<int>();
    NsbGeom &g = c->g;
    g.nx = p->nx; g.ny = p->ny; g.nkx = p->nx / 2 + 1;
    c->N = (size_t)g.nx * g.ny;
    c->M = (size_t)g.nkx * g.ny;
    /* factorisation */
    /* long factors of 1024 (the warp-resident transforms) with short factors up to 32: 32768 = 1024 x 32 */
    int Y1 = p->y1 > 0 ? p->y1 : (g.ny <= 1024 ? g.ny : (g.ny / 32 > 1024 ? g.ny / 32 : 1024));
    int X2 = p->x2 > 0 ? p->x2 : (g.nx <= 1024 ? g.nx : (g.nx / 32 > 1024 ? g.nx / 32 : 1024));
    if (!pow2(Y1) || Y1 > g.ny || Y1 < 2 || !pow2(X2) || X2 > g.nx || X2 < 2) {
        delete c->smem_set; delete c;
        return fail("bad factorisation y1=%d x2=%d", Y1, X2);
    }
    g.Y1 = Y1; g.Y2 = g.ny / Y1; g.X2 = X2; g.X1 = g.nx / X2; g.X2h = X2 / 2;
    if (g.Y2 > 32 || g.X1 > 32) {
        delete c->smem_set; delete c;
        return fail("factorisation needs Y2=%d, X1=%d (max 32 each)", g.Y2, g.X1);
    }
    g.imax = (2 * g.nkx) / 3;  /* spectral_initialization.c:46 */
    g.ncol = g.imax + 1;
    g.ycut = (2 * g.ny) / 6;   /* spectral_initialization.c:48 */
    g.C = p->col_tile > 0 ? p->col_tile : (4096 / Y1 < 4 ? 4 : (4096 / Y1 > 16 ? 16 : 4096 / Y1));
    g.BC = p->beta_batch > 0 ? p->beta_batch : 16;
    const int ncol8 = (g.ncol + 7) & ~7;
    g.PT = ncol8;
    g.PH = (g.X1 == 1) ? ncol8 : g.X2h + 2;
    g.fsT = (size_t)g.ny * g.PT;
    g.fsH = (size_t)g.ny * g.X1 * g.PH;
    g.norm = 1.0 / ((double)g.nx * (double)g.ny);
    c->has_beta = (g.Y2 > 1 || g.X1 > 1);
    c->fast = !(getenv("NSB_DISABLE_FAST") && atoi(getenv("NSB_DISABLE_FAST")));
    c->kW0 = c->kWst = c->kACC = c->kSFD = nullptr;
    c->k_truth = 0;
    c->chZ = c->chP = nullptr;
    c->bufP = nullptr;
    c->tmT_ok = 0;
    c->tmT8_ok = 0;
    c->chCtr = nullptr;
    c->chain_mode = 0;
    c->ch_rows = g.Y2 > NSB_CH_ROWS ? NSB_CH_ROWS_MAX : NSB_CH_ROWS;
    if (c->fast && g.X2 == 1024 && c->has_beta && g.ny % c->ch_rows == 0 && g.Y2 <= c->ch_rows &&
        g.Y1 % (c->ch_rows / g.Y2) == 0) {
        c->chain_mode = getenv("NSB_CHAIN") ? atoi(getenv("NSB_CHAIN")) : 1;
        /* the persistent dataflow kernel keeps a whole 16-column B tile per CTA: 16-point factors only */
        if (c->chain_mode == 2 && (g.Y2 > 16 || g.X1 > 16)) c->chain_mode = 1;
    }
    auto env_int = [](const char *name, int dflt) { return getenv(name) ? atoi(getenv(name)) : dflt; };
    c->ch_slotZ = env_int("NSB_CHAIN_SLOTS_Z", 8);
    c->ch_slotP = env_int("NSB_CHAIN_SLOTS_P", 12);
    c->ch_leadG = env_int("NSB_CHAIN_LEAD_G", 2);
    c->ch_leadF = env_int("NSB_CHAIN_LEAD_F", 8);
    c->ch_ctas = env_int("NSB_CHAIN_CTAS", 2) == 2 ? 2 : 1;
    c->ch_btma = env_int("NSB_CHAIN_BTMA", 1);
    c->ch_bbc = env_int("NSB_CHAIN_BBC", -1);
    c->ch_fbc = env_int("NSB_CHAIN_FBC", 0);
    c->ch_fthreads = env_int("NSB_CHAIN_FTHREADS", 128) == 256 ? 256 : 128; /* F with 32-point factors only */
    c->ch_pieces = env_int("NSB_CHAIN_PIECES", 0);         /* 0 = automatic */
    c->ch_g2 = env_int("NSB_CHAIN_G2", 1);
    c->ch_b512 = env_int("NSB_CHAIN_B512", 1);
    c->ch_ftma = env_int("NSB_CHAIN_FTMA", 1);
    c->tmP_base = nullptr;
    c->ch_gpersist = env_int("NSB_CHAIN_GPERSIST", 0); /* measured: the persistent G is 6 % slower (2.93 vs 2.77 ms) */
    c->alpha_cols = env_int("NSB_ALPHA_COLS", 4) == 8 ? 8 : 4;
    c->alpha_tma = env_int("NSB_ALPHA_TMA", 1);
    c->alpha_persist = env_int("NSB_ALPHA_PERSIST", 1);
    c->alpha_groups = env_int("NSB_ALPHA_GROUPS", 0); /* 0 = automatic: 2 on one device, 1 when slab-decomposed */
    c->tmA_ok = 0;
    if (c->ch_leadG < 1) c->ch_leadG = 1;
    if (c->ch_leadF < 1) c->ch_leadF = 1;
    if (c->ch_slotZ <= c->ch_leadG) c->ch_slotZ = c->ch_leadG + 1;
    if (c->ch_slotP <= c->ch_leadF) c->ch_slotP = c->ch_leadF + 1;
    c->sm_count = 1;
    c->buf_elems = 4 * (g.fsT > g.fsH ? g.fsT : g.fsH);
    /* plain-kernel chain keeps the packed comb spectra Z (2 x ny*X1 x 1024) in bufB */
    if (c->chain_mode && (size_t)2 * g.ny * g.X1 * 1024 > c->buf_elems) c->buf_elems = (size_t)2 * g.ny * g.X1 * 1024;

    CK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, p->device));
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    for (int k = 0; k < 4; ++k) CK(cudaEventCreateWithFlags(&c->ch_ev[k], cudaEventDisableTiming));
    CK(cudaMalloc((void **)&c->bufA, c->buf_elems * sizeof(cd)));
    CK(cudaMalloc((void **)&c->bufB, c->buf_elems * sizeof(cd)));
    CK(cudaMemset(c->bufA, 0, c->buf_elems * sizeof(cd)));
    CK(cudaMemset(c->bufB, 0, c->buf_elems * sizeof(cd)));
    c->bufTp = nullptr;
    if (c->chain_mode) {
        CK(cudaMalloc((void **)&c->bufTp, g.fsT * sizeof(cd)));
        CK(cudaMemset(c->bufTp, 0, g.fsT * sizeof(cd)));
    }
    c->mg_rank = 0; c->mg_world = 1;
    c->mg_cb = g.PT; c->mg_c0 = 0; c->mg_c1 = g.ncol;
    c->mg_rows = g.ny; c->mg_row_shift = 30;
    memset(c->mg_T, 0, sizeof(c->mg_T)); memset(c->mg_Tp, 0, sizeof(c->mg_Tp)); memset(c->mg_sc, 0, sizeof(c->mg_sc));
    c->mg_T[0] = c->bufA; c->mg_Tp[0] = c->bufTp;
    c->mg_ipc = 0; c->mg_epoch = 0; c->mg_nopened = 0;
    CK(cudaEventCreateWithFlags(&c->mg_ev, cudaEventDisableTiming));

    /* wavenumber tables exactly as spectral_initialization.c:25-36 */
    std::vector<double> kx(g.nkx), ky(g.ny);
    for (int i = 0; i < g.nkx; ++i) kx[i] = 2.0 * M_PI * i / p->Lx;
    for (int j = 0; j < g.ny; ++j) ky[j] = (j <= g.ny / 2) ? 2.0 * M_PI * j / p->Ly : 2.0 * M_PI * (j - g.ny) / p->Ly;
    CK(cudaMalloc((void **)&c->d_kx, kx.size() * sizeof(double)));
    CK(cudaMalloc((void **)&c->d_ky, ky.size() * sizeof(double)));
    CK(cudaMemcpy(c->d_kx, kx.data(), kx.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_ky, ky.data(), ky.size() * sizeof(double), cudaMemcpyHostToDevice));
    {
        std::vector<double> kyK(g.ny);
        for (int j2 = 0; j2 < g.Y2; ++j2)
            for (int j1 = 0; j1 < g.Y1; ++j1) kyK[(size_t)j2 * g.Y1 + j1] = ky[(size_t)j1 * g.Y2 + j2];
        CK(cudaMalloc((void **)&c->d_kyK, kyK.size() * sizeof(double)));
        CK(cudaMemcpy(c->d_kyK, kyK.data(), kyK.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    std::vector<cd> w;
    fill_twiddles(w, g.ny);
    CK(cudaMalloc((void **)&c->d_WY, w.size() * sizeof(cd)));
    CK(cudaMemcpy(c->d_WY, w.data(), w.size() * sizeof(cd), cudaMemcpyHostToDevice));
    fill_twiddles(w, g.nx);
    CK(cudaMalloc((void **)&c->d_WX, w.size() * sizeof(cd)));
    CK(cudaMemcpy(c->d_WX, w.data(), w.size() * sizeof(cd), cudaMemcpyHostToDevice));
    g.kx = c->d_kx; g.ky = c->d_ky; g.kyK = c->d_kyK; g.WY = c->d_WY; g.WX = c->d_WX;
    {
        /* the 1024 twiddles of the warp FFT in the order its CTAs stage them (entries of the same
         * long-double tables: bit-identical to gathering them from WY / WX) */
        std::vector<cd> tw(2048), wn;
        for (int pass = 0; pass < 2; ++pass) {
            const int n = pass ? g.nx : g.ny;
            fill_twiddles(wn, n);
            for (int k = 0; k < 1024; ++k) {
                const int r = k >> 5, t = k & 31;
                tw[pass * 1024 + k] = (n >= 1024) ? wn[(size_t)(r * t) * (n / 1024)] : make_double2(1.0, 0.0);
            }
        }
        CK(cudaMalloc((void **)&c->d_TW, tw.size() * sizeof(cd)));
        CK(cudaMemcpy(c->d_TW, tw.data(), tw.size() * sizeof(cd), cudaMemcpyHostToDevice));
        g.TWY = c->d_TW;
        g.TWX = c->d_TW + 1024;
    }

    {
        /* the B kernel's x twiddles in the order its batches consume them (entries of the same table) */
        std::vector<cd> wx, twk;
        fill_twiddles(wx, g.nx);
        const int nk2 = ((g.X2h + 1 + 15) / 16) * 16;
        twk.assign((size_t)nk2 * NSB_TWK_PITCH(g.X1), make_double2(1.0, 0.0));
        for (int k2 = 0; k2 <= g.X2h; ++k2)
            for (int x1 = 0; x1 < g.X1; ++x1) twk[(size_t)k2 * NSB_TWK_PITCH(g.X1) + x1] = wx[((size_t)k2 * x1) % g.nx];
        CK(cudaMalloc((void **)&c->d_TWK, twk.size() * sizeof(cd)));
        CK(cudaMemcpy(c->d_TWK, twk.data(), twk.size() * sizeof(cd), cudaMemcpyHostToDevice));
        g.TWK = c->d_TWK;
    }

    CK(cudaMalloc((void **)&c->d_sc, sizeof(NsbScalars)));
    CK(cudaMemset(c->d_sc, 0, sizeof(NsbScalars)));
    c->mg_sc[0] = c->d_sc;
    CK(cudaMallocHost((void **)&c->h_sc, sizeof(NsbScalars)));
    memset(c->h_sc, 0, sizeof(NsbScalars));
    memset(&c->hv, 0, sizeof(c->hv));
    c->hv.order = 1;
    c->d_force = nullptr; c->force_cap = 0; c->ib = c->jb = 0; c->force_on = 0; c->force_slot = 0;
    for (int s = 0; s < NSB_FORCE_RING; ++s) {
        c->h_force[s] = nullptr; c->h_force_cap[s] = 0;
        CK(cudaEventCreateWithFlags(&c->force_ev[s], cudaEventDisableTiming));
    }
    c->masked_nonzero = 0;
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    void *tmp;
    if (get_field(c, NSB_OMEGA_HAT, &tmp) || get_field(c, NSB_OMEGA_STAGE, &tmp) || get_field(c, NSB_RK_ACC, &tmp)) {
        nsb_ctx_destroy(c);
        return -1;
    }
    CK(cudaStreamSynchronize(c->stream));
    *out = c;
    return 0;
}

extern "C" void nsb_ctx_destroy(nsb_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int f = 0; f < NSB_NUM_FIELDS; ++f) if (c->fields[f]) cudaFree(c->fields[f]);
    cudaFree(c->bufA); cudaFree(c->bufB);
    if (c->kW0) cudaFree(c->kW0);
    if (c->kWst) cudaFree(c->kWst);
    if (c->kACC) cudaFree(c->kACC);
    if (c->kSFD) cudaFree(c->kSFD);
    for (int k = 0; k < c->mg_nopened; ++k) cudaIpcCloseMemHandle(c->mg_opened[k]);
    if (c->bufTp) cudaFree(c->bufTp);
    cudaEventDestroy(c->mg_ev);
    if (c->bufP) cudaFree(c->bufP);
    if (c->chZ) cudaFree(c->chZ);
    if (c->chP) cudaFree(c->chP);
    if (c->chCtr) cudaFree(c->chCtr);
    cudaFree(c->d_kx); cudaFree(c->d_ky); cudaFree(c->d_kyK); cudaFree(c->d_WY); cudaFree(c->d_WX); cudaFree(c->d_TW); cudaFree(c->d_TWK);
    cudaFree(c->d_sc); cudaFreeHost(c->h_sc);
    if (c->d_force) cudaFree(c->d_force);
    for (int s = 0; s < NSB_FORCE_RING; ++s) {
        if (c->h_force[s]) cudaFreeHost(c->h_force[s]);
        cudaEventDestroy(c->force_ev[s]);
    }
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
#ifndef NSB_EMULATE
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
#endif
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->stream2);
    for (int k = 0; k < 4; ++k) cudaEventDestroy(c->ch_ev[k]);
    for (cudaEvent_t e : *c->prof_ev) cudaEventDestroy(e);
    delete c->prof_ev;
    delete c->prof_kind;
    delete c->smem_set;
    delete c;
}

extern "C" int nsb_profile_begin(nsb_ctx *c) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    for (cudaEvent_t e : *c->prof_ev) cudaEventDestroy(e);
    c->prof_ev->clear();
    c->prof_kind->clear();
    c->prof_on = 1;
    return 0;
}

extern "C" int nsb_profile_end(nsb_ctx *c, double ms[6], int count[6]) {
    if (!c || !ms || !count) return fail("nsb_profile_end: null argument");
    CK(cudaSetDevice(c->device));
    c->prof_on = 0;
    CK(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < PK_COUNT; ++k) { ms[k] = 0.0; count[k] = 0; }
    for (size_t k = 0; k < c->prof_kind->size(); ++k) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, (*c->prof_ev)[2 * k], (*c->prof_ev)[2 * k + 1]));
        ms[(*c->prof_kind)[k]] += t;
        count[(*c->prof_kind)[k]] += 1;
    }
    return 0;
}

extern "C" int nsb_sync(nsb_ctx *c) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CKL();
    return 0;
}

extern "C" int nsb_timer_start(nsb_ctx *c) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->ev0, c->stream));
    return 0;
}

extern "C" int nsb_timer_stop(nsb_ctx *c, float *ms) {
    if (!c || !ms) return fail("nsb_timer_stop: null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    CK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    CKL();
    return 0;
}

extern "C" void *nsb_stream(nsb_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" long long nsb_launch_count(nsb_ctx *c) { return c ? c->launches : 0; }

static int pointwise_grid(nsb_ctx *c) {
    size_t blocks = (c->M + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    return (int)blocks;
}

static bool alpha_fast_ok(const nsb_ctx *c) { return c->fast && c->g.Y1 == 1024 && !c->masked_nonzero; }

static int klayout_copy(nsb_ctx *c, bool to_k, cd *ref, cd *K) {
    const NsbGeom &g = c->g;
    const dim3 grid((c->mg_c1 - c->mg_c0 + 31) / 32, g.Y1 / 32, g.Y2), block(32, 8);
    if (to_k) {
        auto k = k_ref_to_klayout;
        NSB_LAUNCH(k, grid, block, 0, c->stream, g, (const cd *)ref, K, c->mg_c0, c->mg_c1);
    } else {
        auto k = k_klayout_to_ref;
        NSB_LAUNCH(k, grid, block, 0, c->stream, g, (const cd *)K, ref, c->mg_c0, c->mg_c1);
    }
    CKL();
    c->launches++;
    return 0;
}

/* make the reference-layout OMEGA_HAT / SFD_HAT current again (after fast-path steps) */
static int sync_ref(nsb_ctx *c) {
    if (!c->k_truth) return 0;
    void *w;
    if (get_field(c, NSB_OMEGA_HAT, &w)) return -1;
    if (klayout_copy(c, false, (cd *)w, c->kW0)) return -1;
    if (c->hv.sfd && c->kSFD) {
        void *s;
        if (get_field(c, NSB_SFD_HAT, &s)) return -1;
        if (klayout_copy(c, false, (cd *)s, c->kSFD)) return -1;
    }
    c->k_truth = 0;
    return 0;
}

/* make the K-layout state current (before fast-path steps) */
static int sync_k(nsb_ctx *c) {
    const size_t bytes = (size_t)(c->mg_c1 - c->mg_c0) * c->g.ny * sizeof(cd);
    if (!c->kW0) {
        CK(cudaMalloc((void **)&c->kW0, bytes));
        CK(cudaMalloc((void **)&c->kWst, bytes));
        CK(cudaMalloc((void **)&c->kACC, bytes));
        CK(cudaMemsetAsync(c->kWst, 0, bytes, c->stream));
        CK(cudaMemsetAsync(c->kACC, 0, bytes, c->stream));
    }
    if (c->hv.sfd && !c->kSFD) CK(cudaMalloc((void **)&c->kSFD, bytes));
    if (c->k_truth) return 0;
    void *w;
    if (get_field(c, NSB_OMEGA_HAT, &w)) return -1;
    if (klayout_copy(c, true, (cd *)w, c->kW0)) return -1;
    if (c->hv.sfd) {
        void *s;
        if (get_field(c, NSB_SFD_HAT, &s)) return -1;
        if (klayout_copy(c, true, (cd *)s, c->kSFD)) return -1;
    }
    c->k_truth = 1;
    return 0;
}

static int scan_masked(nsb_ctx *c) {
    void *w;
    if (get_field(c, NSB_OMEGA_HAT, &w)) return -1;
    CK(cudaMemsetAsync(&c->d_sc->masked_nonzero, 0, sizeof(unsigned), c->stream));
    auto k = k_scan_masked;
    NSB_LAUNCH(k, dim3(pointwise_grid(c)), dim3(256), 0, c->stream, c->g, (const cd *)w, &c->d_sc->masked_nonzero);
    CKL();
    c->launches++;
    unsigned flag = 0;
    CK(cudaMemcpyAsync(&flag, &c->d_sc->masked_nonzero, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->masked_nonzero = flag ? 1 : 0;
    return 0;
}

extern "C" int nsb_upload(nsb_ctx *c, int field, const void *host) {
    if (!c || !host) return fail("nsb_upload: null argument");
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    void *d;
    if (get_field(c, field, &d)) return -1;
    const size_t bytes = field_is_phys(field) ? c->N * sizeof(double) : c->M * sizeof(cd);
    CK(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (field == NSB_OMEGA_HAT) {
        CK(cudaMemsetAsync(&c->d_sc->nonfinite, 0, sizeof(unsigned), c->stream)); /* a new state: new verdict */
        return scan_masked(c);
    }
    return 0;
}

extern "C" int nsb_download(nsb_ctx *c, int field, void *host) {
    if (!c || !host) return fail("nsb_download: null argument");
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    void *d;
    if (get_field(c, field, &d)) return -1;
    const size_t bytes = field_is_phys(field) ? c->N * sizeof(double) : c->M * sizeof(cd);
    CK(cudaMemcpyAsync(host, d, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

/* own-column block [col0, col1) x ny of a spectral field <-> a dense host buffer (per-device restart files of a
 * slab-decomposed run: no gather through the first device) */
extern "C" int nsb_download_columns(nsb_ctx *c, int field, void *host, int *col0, int *col1) {
    if (!c || !host) return fail("nsb_download_columns: null argument");
    if (!field_is_hat(field)) return fail("nsb_download_columns: not a spectral field");
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    void *d;
    if (get_field(c, field, &d)) return -1;
    const int w = c->mg_c1 - c->mg_c0;
    if (col0) *col0 = c->mg_c0;
    if (col1) *col1 = c->mg_c1;
    if (w <= 0) return 0;
    CK(cudaMemcpy2DAsync(host, (size_t)w * sizeof(cd), (const cd *)d + c->mg_c0, (size_t)c->g.nkx * sizeof(cd),
                         (size_t)w * sizeof(cd), (size_t)c->g.ny, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int nsb_zero(nsb_ctx *c, int field) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    void *d;
    if (get_field(c, field, &d)) return -1;
    const size_t bytes = field_is_phys(field) ? c->N * sizeof(double) : c->M * sizeof(cd);
    CK(cudaMemsetAsync(d, 0, bytes, c->stream));
    if (field == NSB_OMEGA_HAT) c->masked_nonzero = 0;
    return 0;
}

extern "C" int nsb_copy(nsb_ctx *c, int dst, int src) {
    if (!c) return fail("null context");
    if (field_is_phys(dst) != field_is_phys(src)) return fail("nsb_copy: field kinds differ");
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    void *d, *s;
    if (get_field(c, dst, &d) || get_field(c, src, &s)) return -1;
    const size_t bytes = field_is_phys(dst) ? c->N * sizeof(double) : c->M * sizeof(cd);
    CK(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, c->stream));
    if (dst == NSB_OMEGA_HAT) return scan_masked(c);
    return 0;
}

extern "C" int nsb_set_hyperviscosity(nsb_ctx *c, const nsb_hyperviscosity *h) {
    if (!c || !h) return fail("nsb_set_hyperviscosity: null argument");
    c->hv.visc = h->enabled ? 1 : 0;
    c->hv.sfd = h->sfd_enabled ? 1 : 0;
    c->hv.order = h->order > 0 ? h->order : 1;
    c->hv.coef = h->coefficient;
    c->hv.kc = h->cutoff_wavenumber;
    c->hv.inv_bw = h->inv_bandwidth;
    c->hv.sfd_strength = h->sfd_strength;
    c->hv.sfd_timescale = h->sfd_timescale;
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    if (c->hv.sfd) {
        void *tmp;
        CK(cudaSetDevice(c->device));
        if (get_field(c, NSB_SFD_HAT, &tmp)) return -1;
    }
    return 0;
}

extern "C" int nsb_set_physics(nsb_ctx *c, double Re, double cfl_factor) {
    if (!c) return fail("null context");
    if (!(Re > 0.0) || !(cfl_factor > 0.0)) return fail("nsb_set_physics: Re and cfl_factor must be positive");
    c->prm.Re = Re;
    c->prm.cfl_factor = cfl_factor;
    return 0;
}

extern "C" int nsb_set_forcing(nsb_ctx *c, int ib, int jb, const double *box) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->device));
    if (!box) { c->force_on = 0; return 0; }
    if (ib < 0 || jb < 0 || ib > c->g.nx / 2 || jb > c->g.ny / 2) return fail("bad forcing box %d x %d", ib, jb);
    const size_t elems = (size_t)(2 * jb + 1) * (ib + 1);
    if (elems > c->force_cap) {
        CK(cudaStreamSynchronize(c->stream));
        if (c->d_force) CK(cudaFree(c->d_force));
        CK(cudaMalloc((void **)&c->d_force, elems * sizeof(cd)));
        c->force_cap = elems;
    }
    const int s = c->force_slot;
    c->force_slot = (s + 1) % NSB_FORCE_RING;
    CK(cudaEventSynchronize(c->force_ev[s])); /* pinned slot free again? */
    if (elems > c->h_force_cap[s]) {
        if (c->h_force[s]) CK(cudaFreeHost(c->h_force[s]));
        CK(cudaMallocHost((void **)&c->h_force[s], elems * sizeof(cd)));
        c->h_force_cap[s] = elems;
    }
    memcpy(c->h_force[s], box, elems * sizeof(cd));
    CK(cudaMemcpyAsync(c->d_force, c->h_force[s], elems * sizeof(cd), cudaMemcpyHostToDevice, c->stream));
    CK(cudaEventRecord(c->force_ev[s], c->stream));
    /* modes of the box that the 2/3 mask removes make the masked region live */
    if (ib > c->g.imax || jb > c->g.ycut) c->masked_nonzero = 1;
    c->ib = ib; c->jb = jb; c->force_on = 1;
    return 0;
}

/* ------------------------------------------------------------ launch helpers */

static int launch_alpha_inv(nsb_ctx *c, int mode, const AlphaInvArgs &a) {
    const NsbGeom &g = c->g;
    const dim3 grid((g.ncol + g.C - 1) / g.C, g.Y2), block(256);
    const size_t smem = (size_t)g.Y1 * g.C * sizeof(cd);
#define NSB_AI(MODE)                                                     \
    {                                                                    \
        auto k = k_alpha_inv<MODE>;                                      \
        if (ensure_smem(c, k, smem)) return -1;                          \
        NSB_LAUNCH(k, grid, block, smem, c->stream, g, a);               \
    }
    prof_pre(c, PK_ALPHA_INV);
    switch (mode) {
        case NSB_SRC_ARRAYS: NSB_AI(NSB_SRC_ARRAYS); break;
        case NSB_SRC_STEP: NSB_AI(NSB_SRC_STEP); break;
        case NSB_SRC_NL: NSB_AI(NSB_SRC_NL); break;
        default: NSB_AI(NSB_SRC_VEL); break;
    }
#undef NSB_AI
    prof_post(c);
    CKL();
    c->launches++;
    return 0;
}

static int launch_alpha_fwd(nsb_ctx *c, int epi, const AlphaFwdArgs &a) {
    const NsbGeom &g = c->g;
    const dim3 grid((g.ncol + g.C - 1) / g.C, g.Y2), block(256);
    const size_t smem = (size_t)g.Y1 * g.C * sizeof(cd);
#define NSB_AF(EPI)                                                      \
    {                                                                    \
        auto k = k_alpha_fwd<EPI>;                                       \
        if (ensure_smem(c, k, smem)) return -1;                          \
        NSB_LAUNCH(k, grid, block, smem, c->stream, g, a);               \
    }
    prof_pre(c, PK_ALPHA_FWD);
    switch (epi) {
        case NSB_EPI_STORE: NSB_AF(NSB_EPI_STORE); break;
        case NSB_EPI_NEG: NSB_AF(NSB_EPI_NEG); break;
        default: NSB_AF(NSB_EPI_RK); break;
    }
#undef NSB_AF
    prof_post(c);
    CKL();
    c->launches++;
    return 0;
}

static size_t beta_smem(const NsbGeom &g) {
    int nd, nm;
    nsb_beta_segs(g.imax, g.X1, g.X2, &nd, &nm);
    return (size_t)g.Y2 * (nd + nm) * g.BC * sizeof(cd);
}

static int beta_batches(const NsbGeom &g) {
    const int nk2 = (g.X1 == 1) ? g.ncol : g.X2h + 1;
    return (nk2 + g.BC - 1) / g.BC;
}

template <int Y2, int X1>
static int launch_beta_t(nsb_ctx *c, bool inverse, const BetaArgs &a) {
    const NsbGeom &g = c->g;
    const size_t smem = beta_smem(g);
    const dim3 block(256);
    prof_pre(c, inverse ? PK_BETA_INV : PK_BETA_FWD);
    if (inverse) {
        auto k = k_beta_inv<Y2, X1>;
        if (ensure_smem(c, k, smem)) return -1;
        NSB_LAUNCH(k, dim3(beta_batches(g), g.Y1, a.nf), block, smem, c->stream, g, a);
    } else {
        auto k = k_beta_fwd<Y2, X1>;
        if (ensure_smem(c, k, smem)) return -1;
        NSB_LAUNCH(k, dim3(beta_batches(g), g.Y1, 1), block, smem, c->stream, g, a);
    }
    prof_post(c);
    CKL();
    c->launches++;
    return 0;
}

template <int Y2>
static int launch_beta_x(nsb_ctx *c, bool inverse, const BetaArgs &a) {
    switch (c->g.X1) {
        case 1: return launch_beta_t<Y2, 1>(c, inverse, a);
        case 2: return launch_beta_t<Y2, 2>(c, inverse, a);
        case 4: return launch_beta_t<Y2, 4>(c, inverse, a);
        case 8: return launch_beta_t<Y2, 8>(c, inverse, a);
        case 16: return launch_beta_t<Y2, 16>(c, inverse, a);
        case 32: return launch_beta_t<Y2, 32>(c, inverse, a);
    }
    return fail("unsupported X1=%d", c->g.X1);
}

static int launch_beta(nsb_ctx *c, bool inverse, const BetaArgs &a) {
    switch (c->g.Y2) {
        case 1: return launch_beta_x<1>(c, inverse, a);
        case 2: return launch_beta_x<2>(c, inverse, a);
        case 4: return launch_beta_x<4>(c, inverse, a);
        case 8: return launch_beta_x<8>(c, inverse, a);
        case 16: return launch_beta_x<16>(c, inverse, a);
        case 32: return launch_beta_x<32>(c, inverse, a);
    }
    return fail("unsupported Y2=%d", c->g.Y2);
}

/* 3-D FP64 tensor map {dim0 doubles, dim1, dim2} with byte strides s1, s2 and a {box0, 1, box2} box,
 * 128-byte swizzle (box0 must be 16 doubles): the tile stores of k_alpha1024<.., TMA> */
static int make_tmap_3d_sw128(nsb_tmap *out, const void *base, size_t dim0, size_t dim1, size_t dim2, size_t s1,
                              size_t s2, int box0, int box2) {
    /* box0 == 16 doubles: 128-byte rows, SWIZZLE_128B; box0 == 8 doubles: 64-byte rows, SWIZZLE_64B */
#ifdef NSB_EMULATE
    out->base = (const double *)base;
    out->dim0 = (long long)dim0; out->dim1 = (long long)dim1; out->dim2 = (long long)dim2;
    out->stride1 = (long long)s1; out->stride2 = (long long)s2;
    out->box0 = box0; out->box1 = 1; out->box2 = box2; out->swz = box0 == 8 ? 2 : 1;
    return 0;
#else
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail("cuTensorMapEncodeTiled is not available");
    const cuuint64_t gdim[3] = {(cuuint64_t)dim0, (cuuint64_t)dim1, (cuuint64_t)dim2};
    const cuuint64_t gstr[2] = {(cuuint64_t)s1, (cuuint64_t)s2};
    const cuuint32_t box[3] = {(cuuint32_t)box0, 1u, (cuuint32_t)box2};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = ((encode_fn)fn)(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(base), gdim, gstr, box,
                                       estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       box0 == 8 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (3-D) failed (%d)", (int)r);
    return 0;
#endif
}

/* T / T' locations and the slab partition of the fused column kernel */
static void fill_alpha_fast(nsb_ctx *c, AlphaFastArgs &fa, const cd *tp, int tp_pitch) {
    const NsbGeom &g = c->g;
    fa.in = tp;
    fa.in_pitch = tp_pitch;
    for (int p = 0; p < NSB_MAXP; ++p) fa.outp[p] = c->mg_T[p < c->mg_world ? p : 0];
    fa.out_fs = (size_t)c->mg_rows * g.PT;
    fa.out_pitch = g.PT;
    fa.row_shift = c->mg_row_shift;
    fa.c0 = c->mg_c0;
    fa.c1 = c->mg_c1;
}

static int launch_alpha1024(nsb_ctx *c, bool fwd, bool inv, const AlphaFastArgs &a_in) {
    const NsbGeom &g = c->g;
    AlphaFastArgs a = a_in;
    const bool tma = c->alpha_tma && g.Y2 > 1 &&
                     (inv || (fwd && c->bufTp && a.in == c->bufTp && a.in_pitch == c->mg_cb));
    if (tma) {
        a.tma_rl = g.Y1 / c->mg_world;
        a.tma_in = (fwd && c->bufTp && a.in == c->bufTp && a.in_pitch == c->mg_cb) ? 1 : 0;
        a.tma_br = a.tma_rl < 256 ? a.tma_rl : 256;
        if (!c->tmA_ok) {
            for (int p = 0; p < c->mg_world; ++p) {
                if (make_tmap_3d_sw128(&c->tmA.m[p], c->mg_T[p], (size_t)2 * g.PT, (size_t)g.Y2, (size_t)4 * a.tma_rl,
                                       (size_t)g.PT * sizeof(cd), (size_t)g.Y2 * g.PT * sizeof(cd), 16, a.tma_br))
                    return -1;
                if (make_tmap_3d_sw128(&c->tmA.m4[p], c->mg_T[p], (size_t)2 * g.PT, (size_t)g.Y2, (size_t)4 * a.tma_rl,
                                       (size_t)g.PT * sizeof(cd), (size_t)g.Y2 * g.PT * sizeof(cd), 8, a.tma_br))
                    return -1;
            }
            if (c->bufTp && make_tmap_3d_sw128(&c->tmA.tp, c->bufTp, (size_t)2 * c->mg_cb, (size_t)g.Y2, (size_t)g.Y1,
                                               (size_t)c->mg_cb * sizeof(cd), (size_t)g.Y2 * c->mg_cb * sizeof(cd), 16, 256))
                return -1;
            if (c->bufTp && make_tmap_3d_sw128(&c->tmA.tp4, c->bufTp, (size_t)2 * c->mg_cb, (size_t)g.Y2, (size_t)g.Y1,
                                               (size_t)c->mg_cb * sizeof(cd), (size_t)g.Y2 * c->mg_cb * sizeof(cd), 8, 256))
                return -1;
            c->tmA_ok = 1;
        }
    }
    prof_pre(c, fwd ? PK_ALPHA_FWD : PK_ALPHA_INV);
#define NSB_A1024(F, I, COLS, T, SMEM)                                                          \
    {                                                                                           \
        auto k = k_alpha1024<F, I, COLS, T>;                                                    \
        const int ntiles = ((a.c1 - a.c0 + COLS - 1) / COLS) * g.Y2;                             \
        /* persistent (one CTA per SM walking the tiles) on one device; slab-decomposed the tile-per-CTA launch is   \
         * faster (measured, 16384^2: 23.2 vs 23.7 ms on 2, 14.2 vs 14.9 on 4, 8.35 vs 8.41 on 8 GPUs: the peer pushes  \
         * of a CTA that holds on to its SM drain behind its own next tile instead of behind another CTA's work) */      \
        const bool persist = (T) && ntiles > c->sm_count && c->alpha_persist && (c->mg_world == 1 || c->alpha_persist > 1); \
        const dim3 grid(persist ? c->sm_count : ntiles), block(COLS * 32);                       \
        if (ensure_smem(c, k, SMEM)) return -1;                                                 \
        NSB_LAUNCH(k, grid, block, SMEM, c->stream, g, a, c->tmA);                              \
    }
    if (tma && (c->alpha_groups == 2 || (c->alpha_groups == 0 && c->mg_world == 1))) {
        /* two independent 4-column groups per CTA: two tiles in flight per SM (measured at 16384^2: 3.95 -> 3.86 ms
         * per fused launch, 2.96 -> 2.76 ms for the stage-1 launch). Slab-decomposed runs keep the 8-column tiles:
         * their 128-byte peer stores beat 64-byte ones by a quarter on NVLink (round 1). */
#define NSB_A1024_2G(F, I)                                                                      \
    {                                                                                           \
        auto k = k_alpha1024<F, I, 4, true, 2>;                                                 \
        const int ntiles = ((a.c1 - a.c0 + 3) / 4) * g.Y2;                                      \
        const bool persist = 2 * c->sm_count < ntiles && c->alpha_persist && (c->mg_world == 1 || c->alpha_persist > 1); \
        const dim3 grid(persist ? c->sm_count : (ntiles + 1) / 2), block(256);                  \
        if (ensure_smem(c, k, NSB_A1024_SMEM_2G)) return -1;                                    \
        NSB_LAUNCH(k, grid, block, NSB_A1024_SMEM_2G, c->stream, g, a, c->tmA);                 \
    }
        if (fwd && inv) NSB_A1024_2G(true, true)
        else if (fwd) NSB_A1024_2G(true, false)
        else NSB_A1024_2G(false, true)
#undef NSB_A1024_2G
    } else if (tma) {
        if (fwd && inv) NSB_A1024(true, true, 8, true, NSB_A1024_SMEM_TMA)
        else if (fwd) NSB_A1024(true, false, 8, true, NSB_A1024_SMEM_TMA)
        else NSB_A1024(false, true, 8, true, NSB_A1024_SMEM_TMA)
    } else if (c->alpha_cols == 8) {
        if (fwd && inv) NSB_A1024(true, true, 8, false, NSB_A1024_SMEM(8))
        else if (fwd) NSB_A1024(true, false, 8, false, NSB_A1024_SMEM(8))
        else NSB_A1024(false, true, 8, false, NSB_A1024_SMEM(8))
    } else {
        if (fwd && inv) NSB_A1024(true, true, 4, false, NSB_A1024_SMEM(4))
        else if (fwd) NSB_A1024(true, false, 4, false, NSB_A1024_SMEM(4))
        else NSB_A1024(false, true, 4, false, NSB_A1024_SMEM(4))
    }
#undef NSB_A1024
    prof_post(c);
    CKL();
    c->launches++;
    return 0;
}

#define NSB_G1024_WARPS 4

static int launch_gamma(nsb_ctx *c, int mode, const GammaArgs &a) {
    const NsbGeom &g = c->g;
    if (mode == NSB_GAMMA_STEP && g.X2 == 1024 && c->fast) {
        GammaFastArgs fa;
        fa.in = a.in; fa.out = a.out; fa.nvalid_in = a.nvalid_in; fa.nvalid_out = a.nvalid_out;
        fa.umax_bits = a.umax_bits;
        fa.npairs = (int)((size_t)g.ny * g.X1 / 2);
        int blocks = (fa.npairs + NSB_G1024_WARPS - 1) / NSB_G1024_WARPS;
        if (blocks > 148 * 4) blocks = 148 * 4;
        const size_t smem = NSB_G1024_SMEM(NSB_G1024_WARPS);
        auto k = k_gamma1024<NSB_G1024_WARPS>;
        if (ensure_smem(c, k, smem)) return -1;
        prof_pre(c, PK_GAMMA);
        NSB_LAUNCH(k, dim3(blocks), dim3(NSB_G1024_WARPS * 32), smem, c->stream, g, fa);
        prof_post(c);
        CKL();
        c->launches++;
        return 0;
    }
    const dim3 grid((unsigned)((size_t)g.ny * g.X1 / 2)), block(256);
    const size_t smem = (size_t)4 * (g.X2 + 1) * sizeof(cd) + 256 * sizeof(double);
#define NSB_GM(MODE)                                                     \
    {                                                                    \
        auto k = k_gamma<MODE>;                                          \
        if (ensure_smem(c, k, smem)) return -1;                          \
        NSB_LAUNCH(k, grid, block, smem, c->stream, g, a);               \
    }
    prof_pre(c, PK_GAMMA);
    switch (mode) {
        case NSB_GAMMA_STEP: NSB_GM(NSB_GAMMA_STEP); break;
        case NSB_GAMMA_C2R: NSB_GM(NSB_GAMMA_C2R); break;
        default: NSB_GM(NSB_GAMMA_R2C); break;
    }
#undef NSB_GM
    prof_post(c);
    CKL();
    c->launches++;
    return 0;
}

/* 2-D FP64 tensor map over a pitched array: dim0 doubles per row, dim1 rows, pitch in bytes,
 * box0 x box1 tile (nsb_platform.h: nsb_tma_2d_g2s). The driver entry point is resolved at
 * run time so the library has no link-time dependency on libcuda. */
static int make_tmap_2d(nsb_tmap *out, const void *base, size_t dim0, size_t dim1, size_t pitch_bytes, int box0,
                        int box1) {
#ifdef NSB_EMULATE
    out->base = (const double *)base;
    out->dim0 = (long long)dim0;
    out->dim1 = (long long)dim1;
    out->stride1 = (long long)pitch_bytes;
    out->box0 = box0;
    out->box1 = box1;
    return 0;
#else
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return fail("cuTensorMapEncodeTiled is not available");
        encode = (encode_fn)fn;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
    const cuuint64_t gstr[1] = {(cuuint64_t)pitch_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(base), gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
#endif
}

/* T (bufA, 4 fields, own rows) -> T' (bufTp of the column owners): B, G, F of nsb_chain.cuh */
static int launch_chain(nsb_ctx *c, bool want_umax) {
    const NsbGeom &g = c->g;
    ChainArgs a;
    memset(&a, 0, sizeof(a));
    a.T = c->bufA;
    a.rowsT = c->mg_rows;
    a.ybase = c->mg_rank * c->mg_rows;
    for (int p = 0; p < NSB_MAXP; ++p) a.Tpp[p] = c->mg_Tp[p < c->mg_world ? p : 0];
    a.cb = c->mg_cb;
    const int R = c->ch_rows;
    const bool big = g.Y2 > 16 || g.X1 > 16;
    a.R = R;
    a.nslab = c->mg_rows / R;
    a.GY = R / g.Y2;
    a.NQ = R * g.X1;
    a.nbatch = (g.X2h + 1 + g.BC - 1) / g.BC; /* every k2 of a comb spectrum is written, also when X1 == 1 */
    nsb_beta_segs(g.imax, g.X1, g.X2, &a.nd, &a.nm);
    a.umax_bits = &c->d_sc->umax_bits;
    const int NP = a.NQ / 2;
    const int W = (c->chain_mode == 2 && c->ch_ctas == 2) ? (want_umax ? 3 : 4) : (want_umax ? 6 : 8);
    const size_t tileB = (size_t)2 * R * (a.nd + a.nm) * g.BC * sizeof(cd);
    const size_t twtab = (size_t)g.BC * NSB_TWK_PITCH(g.X1) * sizeof(cd);
    /* F keeps one tile per CTA; with 32-point factors a 16-column tile (180 KB) would leave one CTA per
     * SM, so its batches narrow to 8 columns there (16384^2 and below: unchanged, 45 KB) */
    int bcF = g.BC;
    while (bcF > 4 && (tileB / 2 + twtab) / (size_t)(g.BC / bcF) > 100 * 1024) bcF /= 2;
    if (c->ch_fbc > 0 && c->ch_fbc <= g.BC && g.BC % c->ch_fbc == 0) bcF = c->ch_fbc; /* NSB_CHAIN_FBC */
    const size_t tileF = (tileB / 2 + twtab) / (size_t)(g.BC / bcF);
    NsbGeom gF = g;
    gF.BC = bcF;
    const size_t smemG = NSB_CH_TW_BYTES + (size_t)W * (want_umax ? NSB_CHG_WARP_BYTES(1) : NSB_CHG_WARP_BYTES(0));
    if (c->chain_mode == 1) {
        /* Z in bufB, P in bufA (T is dead once B has run), T' into the owners' exchange buffers */
        const int total_slabs = a.nslab;
        a.nslotZ = a.nslotP = total_slabs;
        a.Z = c->bufB;
        a.P = c->bufA;
        /* Pieces: the slabs are independent through B, G and F, so the chain can run as `pieces` slab
         * ranges alternating between two streams, staggered so that the memory-bound B of one range
         * shares the SMs with the FP64-bound G of the previous one. That needs CTAs small enough to
         * co-reside: B with 8-column batches (128 threads, 2 x 45 KB) and G with 4 warps (112 KB). */
        /* measured at 16384^2 (1024 slabs): 16 ranges 48.7 ms/step against 50.3 for one; at 4096^2 (256
         * slabs) every split is slower. True co-residency (one B CTA + one G CTA per SM, also with a
         * 16-column one-stage B) was slower still, 54-58 ms: B needs all eight warps of an SM; the gain
         * comes from overlapping the kernels' tails. */
        /* slab-decomposed, ranges of 32 slabs, measured at 16384^2: 8 devices (128 slabs each) 1 / 2 / 4 / 8 /
         * 16 ranges 8.93 / 8.83 / 8.70 / 9.01 / 9.23 ms per step; 4 devices (256 slabs each) 1 / 4 / 8 ranges
         * 15.55 / 15.48 / 15.18 ms */
        const int auto_pieces =
            total_slabs >= 512 ? ((c->ch_b512 && g.X1 >= 2 && !big) ? 4 : 16)
                               : ((c->mg_world > 1 && !big && total_slabs >= 128) ? total_slabs / 32 : 1);
        int pieces = c->prof_on ? 1 : (c->ch_pieces > 0 ? c->ch_pieces : auto_pieces);
        if (pieces > total_slabs) pieces = total_slabs;
        const bool split = pieces > 1;
        if (split) {
            /* P may no longer live in bufA: G of one range writes it while B of the next still reads T there */
            if (!c->bufP) CK(cudaMalloc((void **)&c->bufP, (size_t)total_slabs * NP * 1024 * sizeof(cd)));
            a.P = c->bufP;
        }
        const int Wg = split ? (want_umax ? 3 : 4) : W;
        const size_t smemGp = NSB_CH_TW_BYTES + (size_t)Wg * (want_umax ? NSB_CHG_WARP_BYTES(1) : NSB_CHG_WARP_BYTES(0));
        /* measured: 8-column batches win while the tile is small (4096^2: 0.31 vs 0.40 ms), 16-column
         * batches with one 256-thread CTA per SM win at 16384^2 (3.56 vs 5.58 ms) when B runs alone */
        /* round 2: with the split step 2 (512 threads, 128 registers) the 16-column batch wins at 16384^2 also
         * inside slab ranges (43.0 ms per step in 4 ranges against 45.9 with 8-column batches in 16) */
        const bool b512 = c->ch_b512 && g.X1 >= 2 && !big;
        bool half = g.BC == 16 && (c->ch_bbc == 8 || (c->ch_bbc < 0 && ((split && !b512) || tileB <= 48 * 1024)));
        /* two tile stages must fit one SM: 32-point factors (32 rows x up to 22 segments per column) leave
         * room for 8- or 4-column batches only; such CTAs are alone on their SM (256 threads) */
        int bdiv = half ? 2 : 1;
        const size_t smem_cap = 227 * 1024;
        while (g.BC / bdiv > 4 && 2 * (tileB / bdiv) + 64 + 128 + 2 * twtab / bdiv > smem_cap) bdiv *= 2;
        const bool narrow = bdiv > 1;
        const bool two_ctas = half && bdiv == 2 && 2 * (2 * (tileB / 2) + 64 + 128 + 2 * twtab / 2 + 1024) <= smem_cap;
        half = narrow;
        const bool use_tma = c->ch_btma && 2 * (tileB / bdiv) + 64 + 128 + 2 * twtab / bdiv <= smem_cap && 2 * g.BC <= 256;
        if (!use_tma && tileB > smem_cap)
            return fail("the chain's B tile (%zu bytes) needs the TMA-fed kernel at this size (NSB_CHAIN_BTMA=1)", tileB);
        NsbGeom gB = g;
        ChainArgs aB = a;
        if (use_tma && narrow) {
            gB.BC = g.BC / bdiv;
            aB.nbatch = (g.X2h + 1 + gB.BC - 1) / gB.BC;
        }
        aB.fd_perg = nsb_fastdiv_make((unsigned)((aB.nd + aB.nm) * gB.BC));
        aB.fd_nb = nsb_fastdiv_make((unsigned)aB.nbatch);
        aB.fd_nb2 = nsb_fastdiv_make((unsigned)(2 * aB.nbatch));
        nsb_tmap *tm = half ? &c->tmT8 : &c->tmT;
        if (use_tma) {
            int *tm_ok = half ? &c->tmT8_ok : &c->tmT_ok;
            if (!*tm_ok) {
                if (make_tmap_2d(tm, c->bufA, (size_t)2 * g.PT, (size_t)4 * c->mg_rows, (size_t)g.PT * sizeof(cd),
                                 2 * gB.BC, R))
                    return -1;
                *tm_ok = 1;
            }
        }
        const size_t tB = tileB / bdiv;
        const size_t smemB = 2 * tB + 64 + 128 + 2 * twtab / bdiv;
        if (split) {
            CK(cudaEventRecord(c->ch_ev[0], c->stream)); /* everything before the chain (alpha, barrier) */
            CK(cudaStreamWaitEvent(c->stream2, c->ch_ev[0], 0));
        }
        for (int pc = 0; pc < pieces; ++pc) {
            cudaStream_t st = (pc & 1) ? c->stream2 : c->stream;
            const int s0 = (int)((long long)total_slabs * pc / pieces), s1 = (int)((long long)total_slabs * (pc + 1) / pieces);
            a.slab0 = aB.slab0 = s0;
            a.nslab = aB.nslab = s1 - s0;
            /* stagger: B of this piece starts when B of the previous piece is done */
            if (split && pc > 0) CK(cudaStreamWaitEvent(st, c->ch_ev[1 + ((pc - 1) & 1)], 0));
            if (use_tma) {
                prof_pre(c, PK_BETA_INV);
                if (big) { /* 32-point factors: own instantiation, so the kernels below keep their registers */
                    auto k = k_chain_B_tma<256, 1, true>;
                    if (ensure_smem(c, k, smemB, true)) return -1;
                    NSB_LAUNCH(k, dim3(c->sm_count), dim3(256), smemB, st, gB, aB, *tm);
                } else if (c->ch_b512 && g.X1 >= 2 && !narrow) {
                    auto k = k_chain_B_tma<512, 1, false, true>;
                    if (ensure_smem(c, k, smemB, true)) return -1;
                    NSB_LAUNCH(k, dim3(c->sm_count), dim3(512), smemB, st, gB, aB, *tm);
                } else if (c->ch_b512 && g.X1 >= 2 && two_ctas) {
                    auto k = k_chain_B_tma<256, 2, false, true>;
                    if (ensure_smem(c, k, smemB, true)) return -1;
                    NSB_LAUNCH(k, dim3(2 * c->sm_count), dim3(256), smemB, st, gB, aB, *tm);
                } else if (two_ctas) {
                    auto k = k_chain_B_tma<128, 2, false>;
                    if (ensure_smem(c, k, smemB, true)) return -1;
                    NSB_LAUNCH(k, dim3(2 * c->sm_count), dim3(128), smemB, st, gB, aB, *tm);
                } else {
                    auto k = k_chain_B_tma<256, 1, false>;
                    if (ensure_smem(c, k, smemB, true)) return -1;
                    NSB_LAUNCH(k, dim3(c->sm_count), dim3(256), smemB, st, gB, aB, *tm);
                }
                prof_post(c);
            } else {
                prof_pre(c, PK_BETA_INV);
                if (big) {
                    auto k = k_chain_B<true>;
                    if (ensure_smem(c, k, tileB, true)) return -1;
                    NSB_LAUNCH(k, dim3(a.nbatch, 2, a.nslab), dim3(256), tileB, st, g, a);
                } else {
                    auto k = k_chain_B<false>;
                    if (ensure_smem(c, k, tileB, true)) return -1;
                    NSB_LAUNCH(k, dim3(a.nbatch, 2, a.nslab), dim3(256), tileB, st, g, a);
                }
                prof_post(c);
            }
            CKL();
            c->launches++;
            if (split) CK(cudaEventRecord(c->ch_ev[1 + (pc & 1)], st));
            const dim3 gridG((NP + Wg - 1) / Wg, a.nslab);
            prof_pre(c, PK_GAMMA);
            if (c->ch_g2) {
                /* persistent, 8 warps per SM in both variants (parking in tensor memory) */
                if (want_umax) {
                    auto k = k_chain_G2<true>;
                    if (ensure_smem(c, k, NSB_CHG2_SMEM, true)) return -1;
                    NSB_LAUNCH(k, dim3(c->sm_count), dim3(256), NSB_CHG2_SMEM, st, g, a);
                } else {
                    auto k = k_chain_G2<false>;
                    if (ensure_smem(c, k, NSB_CHG2_SMEM, true)) return -1;
                    NSB_LAUNCH(k, dim3(c->sm_count), dim3(256), NSB_CHG2_SMEM, st, g, a);
                }
            } else if (c->ch_gpersist && !split) {
                if (want_umax) {
                    auto k = k_chain_Gp<true>;
                    if (ensure_smem(c, k, smemGp, true)) return -1;
                    NSB_LAUNCH(k, dim3(c->sm_count), dim3(Wg * 32), smemGp, st, g, a);
                } else {
                    auto k = k_chain_Gp<false>;
                    if (ensure_smem(c, k, smemGp, true)) return -1;
                    NSB_LAUNCH(k, dim3(c->sm_count), dim3(Wg * 32), smemGp, st, g, a);
                }
            } else if (want_umax) {
                auto k = k_chain_G<true>;
                if (ensure_smem(c, k, smemGp, true)) return -1;
                NSB_LAUNCH(k, gridG, dim3(Wg * 32), smemGp, st, g, a);
            } else {
                auto k = k_chain_G<false>;
                if (ensure_smem(c, k, smemGp, true)) return -1;
                NSB_LAUNCH(k, gridG, dim3(Wg * 32), smemGp, st, g, a);
            }
            prof_post(c);
            CKL();
            c->launches++;
            {
                prof_pre(c, PK_BETA_FWD);
                ChainArgs aF = a;
                aF.nbatch = (g.X2h + 1 + bcF - 1) / bcF;
                if (big) {
                    /* 182 registers (one 32-point array per thread): 128-thread CTAs so that two of them,
                     * with their 90 KB tiles, share an SM and overlap each other's load and barrier phases */
                    auto k = k_chain_F<true>;
                    if (ensure_smem(c, k, tileF, true)) return -1;
                    NSB_LAUNCH(k, dim3(aF.nbatch, a.nslab), dim3(c->ch_fthreads), tileF, st, gF, aF);
                } else if (c->ch_ftma && g.X1 >= 2 && NP <= 256 && bcF == g.BC &&
                           4 * (size_t)NP * g.BC * sizeof(cd) + tileB / 2 + 2 * twtab + 256 <= 227 * 1024) {
                    /* persistent, P boxes by TMA, double buffered */
                    if (c->tmP_base != (const void *)a.P) {
                        if (make_tmap_2d(&c->tmP, a.P, 2048, (size_t)total_slabs * NP, 1024 * sizeof(cd), 2 * g.BC, NP))
                            return -1;
                        c->tmP_base = (const void *)a.P;
                    }
                    aF.fd_nb = nsb_fastdiv_make((unsigned)aF.nbatch);
                    const size_t smemF = 4 * (size_t)NP * g.BC * sizeof(cd) + tileB / 2 + 2 * twtab + 256;
                    auto k = k_chain_F_tma;
                    if (ensure_smem(c, k, smemF, true)) return -1;
                    NSB_LAUNCH(k, dim3(c->sm_count), dim3(256), smemF, st, g, aF, c->tmP);
                } else {
                    auto k = k_chain_F<false>;
                    if (ensure_smem(c, k, tileF, true)) return -1;
                    NSB_LAUNCH(k, dim3(aF.nbatch, a.nslab), dim3(256), tileF, st, gF, aF);
                }
                prof_post(c);
                CKL();
                c->launches++;
            }
        }
        if (split) { /* the main stream continues when the other one has drained */
            CK(cudaEventRecord(c->ch_ev[3], c->stream2));
            CK(cudaStreamWaitEvent(c->stream, c->ch_ev[3], 0));
        }
        return 0;
    }
    /* persistent dataflow kernel over L2-resident rings */
    a.nslotZ = c->ch_slotZ < a.nslab ? c->ch_slotZ : a.nslab;
    a.nslotP = c->ch_slotP < a.nslab ? c->ch_slotP : a.nslab;
    a.leadG = c->ch_leadG;
    a.leadF = c->ch_leadF;
    if (a.nslotZ <= a.leadG) a.leadG = a.nslotZ > 1 ? a.nslotZ - 1 : 1;
    if (a.nslotP <= a.leadF) a.leadF = a.nslotP > 1 ? a.nslotP - 1 : 1;
    if (a.nslotZ <= a.leadG || a.nslotP <= a.leadF) { /* single-slab grids: no ring reuse can occur */
        a.nslotZ = a.leadG + 1;
        a.nslotP = a.leadF + 1;
    }
    const size_t ctr_n = 1 + 3 * (size_t)a.nslab;
    if (!c->chZ) {
        CK(cudaMalloc((void **)&c->chZ, (size_t)(c->ch_slotZ + 1) * 2 * a.NQ * 1024 * sizeof(cd)));
        CK(cudaMalloc((void **)&c->chP, (size_t)(c->ch_slotP + 1) * NP * 1024 * sizeof(cd)));
        CK(cudaMalloc((void **)&c->chCtr, ctr_n * sizeof(unsigned)));
    }
    a.Z = c->chZ;
    a.P = c->chP;
    a.ctr = c->chCtr;
    CK(cudaMemsetAsync(c->chCtr, 0, ctr_n * sizeof(unsigned), c->stream));
    size_t smem = smemG;
    if (NSB_CH_TW_BYTES + tileB > smem) smem = NSB_CH_TW_BYTES + tileB;
    if (NSB_CH_TW_BYTES + tileF > smem) smem = NSB_CH_TW_BYTES + tileF;
    prof_pre(c, PK_GAMMA);
#define NSB_CHAIN_LAUNCH(UM, WW, MB)                                                            \
    {                                                                                           \
        auto k = k_chain<UM, WW, MB>;                                                           \
        if (ensure_smem(c, k, smem, true)) return -1;                                           \
        NSB_LAUNCH(k, dim3(c->sm_count * MB), dim3(WW * 32), smem, c->stream, g, a);            \
    }
    if (c->ch_ctas == 2) {
        if (want_umax) NSB_CHAIN_LAUNCH(true, 3, 2) else NSB_CHAIN_LAUNCH(false, 4, 2)
    } else {
        if (want_umax) NSB_CHAIN_LAUNCH(true, 6, 1) else NSB_CHAIN_LAUNCH(false, 8, 1)
    }
#undef NSB_CHAIN_LAUNCH
    prof_post(c);
    CKL();
    c->launches++;
    return 0;
}

/* spectral (masked) -> physical for nf fields produced by alpha source `mode` */
static int inverse_fields(nsb_ctx *c, int mode, int nf, const cd *const *src, double *const *phys_out,
                          bool want_umax, bool want_sums) {
    const NsbGeom &g = c->g;
    AlphaInvArgs ai;
    memset(&ai, 0, sizeof(ai));
    ai.nf = nf;
    for (int f = 0; f < 4; ++f) ai.src[f] = src[f];
    ai.out = c->bufA; ai.out_fs = g.fsT; ai.out_pitch = g.PT;
    if (launch_alpha_inv(c, mode, ai)) return -1;
    const cd *H = c->bufA;
    if (c->has_beta) {
        BetaArgs b; b.in = c->bufA; b.out = c->bufB; b.nf = nf;
        if (launch_beta(c, true, b)) return -1;
        H = c->bufB;
    }
    GammaArgs ga;
    memset(&ga, 0, sizeof(ga));
    ga.in = H;
    ga.nvalid_in = (g.X1 == 1) ? g.ncol : g.X2h + 1;
    ga.nf = nf;
    for (int f = 0; f < 4; ++f) ga.phys_out[f] = (f < nf) ? phys_out[f] : nullptr;
    ga.umax_bits = want_umax ? &c->d_sc->umax_bits : nullptr;
    ga.sums = want_sums ? c->d_sc->sums : nullptr;
    return launch_gamma(c, NSB_GAMMA_C2R, ga);
}

/* J-bearing H' (in gamma_out) -> spectral, epilogue `epi` */
static int forward_tail(nsb_ctx *c, const cd *Hp, int epi, AlphaFwdArgs &af) {
    const NsbGeom &g = c->g;
    const cd *Tp = Hp;
    if (c->has_beta) {
        cd *dst = (Hp == c->bufA) ? c->bufB : c->bufA;
        BetaArgs b; b.in = Hp; b.out = dst; b.nf = 1;
        if (launch_beta(c, false, b)) return -1;
        Tp = dst;
    }
    af.in = Tp;
    af.in_pitch = g.PT;
    return launch_alpha_fwd(c, epi, af);
}

/* 4 derivative fields -> Jacobian -> forward, one RK-stage worth of transforms */
static int nonlinear_pipeline(nsb_ctx *c, int src_mode, const cd *s0, const cd *s1, bool want_umax, int epi,
                              AlphaFwdArgs &af) {
    const NsbGeom &g = c->g;
    AlphaInvArgs ai;
    memset(&ai, 0, sizeof(ai));
    ai.nf = 4; ai.src[0] = s0; ai.src[1] = s1;
    ai.out = c->bufA; ai.out_fs = g.fsT; ai.out_pitch = g.PT;
    if (launch_alpha_inv(c, src_mode, ai)) return -1;
    const cd *H = c->bufA;
    cd *Hp = c->bufB;
    if (c->has_beta) {
        BetaArgs b; b.in = c->bufA; b.out = c->bufB; b.nf = 4;
        if (launch_beta(c, true, b)) return -1;
        H = c->bufB; Hp = c->bufA;
    }
    GammaArgs ga;
    memset(&ga, 0, sizeof(ga));
    ga.in = H;
    ga.nvalid_in = (g.X1 == 1) ? g.ncol : g.X2h + 1;
    ga.out = Hp;
    ga.nvalid_out = ga.nvalid_in;
    ga.nf = 4;
    ga.umax_bits = want_umax ? &c->d_sc->umax_bits : nullptr;
    if (launch_gamma(c, NSB_GAMMA_STEP, ga)) return -1;
    return forward_tail(c, Hp, epi, af);
}

/* --------------------------------------------------------------- public ops */

extern "C" int nsb_physical_to_spectral(nsb_ctx *c, int src_phys, int dst_hat) {
    if (!c) return fail("null context");
    if (!field_is_phys(src_phys) || !field_is_hat(dst_hat)) return fail("nsb_physical_to_spectral: bad fields");
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    void *src, *dst;
    if (get_field(c, src_phys, &src) || get_field(c, dst_hat, &dst)) return -1;
    const NsbGeom &g = c->g;
    GammaArgs ga;
    memset(&ga, 0, sizeof(ga));
    ga.phys_in = (const double *)src;
    ga.out = c->bufA;
    ga.nvalid_out = (g.X1 == 1) ? g.ncol : g.X2h + 1;
    if (launch_gamma(c, NSB_GAMMA_R2C, ga)) return -1;
    CK(cudaMemsetAsync(dst, 0, c->M * sizeof(cd), c->stream));
    AlphaFwdArgs af;
    memset(&af, 0, sizeof(af));
    af.out = (cd *)dst;
    if (forward_tail(c, c->bufA, NSB_EPI_STORE, af)) return -1;
    if (dst_hat == NSB_OMEGA_HAT) c->masked_nonzero = 0;
    return 0;
}

extern "C" int nsb_spectral_to_physical(nsb_ctx *c, int nf, const int *src_hat, const int *dst_phys) {
    if (!c || !src_hat || !dst_phys) return fail("nsb_spectral_to_physical: null argument");
    if (nf < 1 || nf > 4) return fail("nsb_spectral_to_physical: nf must be 1..4");
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    const cd *src[4] = {nullptr, nullptr, nullptr, nullptr};
    double *dst[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int f = 0; f < nf; ++f) {
        void *s, *d;
        if (!field_is_hat(src_hat[f]) || !field_is_phys(dst_phys[f])) return fail("nsb_spectral_to_physical: bad fields");
        if (get_field(c, src_hat[f], &s) || get_field(c, dst_phys[f], &d)) return -1;
        src[f] = (const cd *)s; dst[f] = (double *)d;
    }
    return inverse_fields(c, NSB_SRC_ARRAYS, nf, src, dst, false, false);
}

extern "C" int nsb_apply_dealiasing(nsb_ctx *c, int field_hat) {
    if (!c) return fail("null context");
    if (!field_is_hat(field_hat)) return fail("nsb_apply_dealiasing: bad field");
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    void *f;
    if (get_field(c, field_hat, &f)) return -1;
    auto k = k_apply_dealiasing;
    NSB_LAUNCH(k, dim3(pointwise_grid(c)), dim3(256), 0, c->stream, c->g, (cd *)f);
    CKL();
    c->launches++;
    if (field_hat == NSB_OMEGA_HAT) c->masked_nonzero = 0;
    return 0;
}

extern "C" int nsb_velocity_from_vorticity(nsb_ctx *c) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    void *w, *psi, *uh, *vh;
    if (get_field(c, NSB_OMEGA_HAT, &w) || get_field(c, NSB_PSI_HAT, &psi) || get_field(c, NSB_U_HAT, &uh) ||
        get_field(c, NSB_V_HAT, &vh))
        return -1;
    auto k = k_velocity_from_vorticity;
    NSB_LAUNCH(k, dim3(pointwise_grid(c)), dim3(256), 0, c->stream, c->g, (const cd *)w, (cd *)psi, (cd *)uh, (cd *)vh);
    CKL();
    c->launches++;
    return 0;
}

extern "C" int nsb_compute_derivatives(nsb_ctx *c) {
    if (nsb_velocity_from_vorticity(c)) return -1;
    void *w, *u, *v, *om;
    if (get_field(c, NSB_OMEGA_HAT, &w) || get_field(c, NSB_U, &u) || get_field(c, NSB_V, &v) ||
        get_field(c, NSB_OMEGA, &om))
        return -1;
    const cd *src[4] = {(const cd *)w, nullptr, nullptr, nullptr};
    double *dst[4] = {(double *)u, (double *)v, (double *)om, nullptr};
    return inverse_fields(c, NSB_SRC_VEL, 3, src, dst, false, false);
}

extern "C" int nsb_nonlinear_terms(nsb_ctx *c) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    void *psi, *w, *nl;
    if (get_field(c, NSB_PSI_HAT, &psi) || get_field(c, NSB_OMEGA_HAT, &w) || get_field(c, NSB_NONLINEAR_HAT, &nl))
        return -1;
    CK(cudaMemsetAsync(nl, 0, c->M * sizeof(cd), c->stream));
    AlphaFwdArgs af;
    memset(&af, 0, sizeof(af));
    af.out = (cd *)nl;
    return nonlinear_pipeline(c, NSB_SRC_NL, (const cd *)psi, (const cd *)w, false, NSB_EPI_NEG, af);
}

extern "C" int nsb_hyperviscosity_apply(nsb_ctx *c, int rhs_hat, int state_hat) {
    if (!c) return fail("null context");
    if (!field_is_hat(rhs_hat) || !field_is_hat(state_hat)) return fail("nsb_hyperviscosity_apply: bad fields");
    if (!c->hv.visc && !c->hv.sfd) return 0; /* hyperviscosity.c:120-122 */
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    void *rhs, *st, *sfd = nullptr;
    if (get_field(c, rhs_hat, &rhs) || get_field(c, state_hat, &st)) return -1;
    if (c->hv.sfd && get_field(c, NSB_SFD_HAT, &sfd)) return -1;
    auto k = k_hv_apply;
    NSB_LAUNCH(k, dim3(pointwise_grid(c)), dim3(256), 0, c->stream, c->g, c->hv, (cd *)rhs, (const cd *)st, (const cd *)sfd);
    CKL();
    c->launches++;
    return 0;
}

extern "C" int nsb_hyperviscosity_post_step(nsb_ctx *c, int state_hat, double dt) {
    if (!c) return fail("null context");
    if (!field_is_hat(state_hat)) return fail("nsb_hyperviscosity_post_step: bad field");
    if (!(dt > 0.0) || !c->hv.sfd || !(c->hv.sfd_strength > 0.0) || !(c->hv.sfd_timescale > 0.0)) return 0;
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    void *st, *sfd;
    if (get_field(c, state_hat, &st) || get_field(c, NSB_SFD_HAT, &sfd)) return -1;
    double alpha = dt / c->hv.sfd_timescale;
    alpha = alpha > 1.0 ? 1.0 : (alpha < 0.0 ? 0.0 : alpha);
    auto k = k_sfd_post;
    NSB_LAUNCH(k, dim3(pointwise_grid(c)), dim3(256), 0, c->stream, c->g, alpha, (cd *)sfd, (const cd *)st);
    CKL();
    c->launches++;
    return 0;
}

static int launch_cfl(nsb_ctx *c, bool local = false) {
    const double dx_min = fmin(c->prm.Lx / c->g.nx, c->prm.Ly / c->g.ny);
    if (c->mg_world > 1 && !local) {
        /* max |u|^2 over the slabs of all devices, read from their scalar blocks (peer memory) */
        NsbPeerScalars ps;
        for (int p = 0; p < NSB_MAXP; ++p) ps.sc[p] = c->mg_sc[p < c->mg_world ? p : 0];
        auto k = k_cfl_dt_multi;
        NSB_LAUNCH(k, dim3(1), dim3(32), 0, c->stream, c->d_sc, ps, c->mg_world, c->prm.cfl_factor, dx_min, c->prm.Re,
                   NSB_MIN_DT, NSB_MAX_DT);
    } else {
        auto k = k_cfl_dt;
        NSB_LAUNCH(k, dim3(1), dim3(32), 0, c->stream, c->d_sc, c->prm.cfl_factor, dx_min, c->prm.Re, NSB_MIN_DT,
                   NSB_MAX_DT);
    }
    CKL();
    c->launches++;
    return 0;
}

static int fetch_scalars(nsb_ctx *c) {
    CK(cudaMemcpyAsync(c->h_sc, c->d_sc, sizeof(NsbScalars), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int nsb_cfl_timestep_uv(nsb_ctx *c, double *dt_out) {
    if (!c || !dt_out) return fail("nsb_cfl_timestep_uv: null argument");
    CK(cudaSetDevice(c->device));
    void *u, *v;
    if (get_field(c, NSB_U, &u) || get_field(c, NSB_V, &v)) return -1;
    CK(cudaMemsetAsync(&c->d_sc->umax_bits, 0, sizeof(unsigned long long), c->stream));
    size_t blocks = (c->N + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    auto k = k_max_speed2;
    NSB_LAUNCH(k, dim3((unsigned)blocks), dim3(256), 256 * sizeof(double), c->stream, (const double *)u,
               (const double *)v, c->N, &c->d_sc->umax_bits);
    CKL();
    c->launches++;
    if (launch_cfl(c, true) || fetch_scalars(c)) return -1; /* whole field on this device: no peer maxima */
    *dt_out = c->h_sc->dt;
    return 0;
}

extern "C" int nsb_cfl_timestep(nsb_ctx *c, double *dt_out) {
    if (!c || !dt_out) return fail("nsb_cfl_timestep: null argument");
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    void *w;
    if (get_field(c, NSB_OMEGA_HAT, &w)) return -1;
    CK(cudaMemsetAsync(&c->d_sc->umax_bits, 0, sizeof(unsigned long long), c->stream));
    const cd *src[4] = {(const cd *)w, nullptr, nullptr, nullptr};
    double *dst[4] = {nullptr, nullptr, nullptr, nullptr};
    if (inverse_fields(c, NSB_SRC_VEL, 2, src, dst, true, false)) return -1;
    if (launch_cfl(c, true) || fetch_scalars(c)) return -1;
    *dt_out = c->h_sc->dt;
    return 0;
}

/*
 * The slab-decomposed fast path (Y1 == X2 == 1024, chain kernels) as a sequence of phases
 * with a device barrier between them:
 *   0           K-layout state current, dt, alpha_inv of stage 1     -> T pushed to row owners
 *   1,3,5,7     chain B/G/F of stage 1..4                            -> T' pushed to column owners
 *   2,4,6       (stage 1: CFL dt from the max over all slabs) alpha fwd + RK + inv of the next stage
 *   8           alpha fwd + RK final combine, t += dt
 */
#define NSB_FAST_PHASES 9

static void fill_rk(nsb_ctx *c, NsbRk &rk, int stage) {
    memset(&rk, 0, sizeof(rk));
    rk.stage = stage;
    rk.W0 = c->kW0; rk.Wst = c->kWst; rk.ACC = c->kACC; rk.SFD = c->kSFD;
    rk.dt = &c->d_sc->dt;
    rk.inv_Re = 1.0 / c->prm.Re;
    rk.hv = c->hv;
    rk.FB = c->force_on ? c->d_force : nullptr;
    rk.ib = c->ib; rk.jb = c->jb;
}

static int fast_phase(nsb_ctx *c, int ph, double dt) {
    static const char *const names[NSB_FAST_PHASES] = {
        "nsb stage 1: spectral ops + inverse y FFT (k_alpha1024)", "nsb stage 1: chain B/G/F (x transforms, Jacobian, max|u|)",
        "nsb stage 1: CFL dt, forward y FFT + RK + stage 2 inverse", "nsb stage 2: chain B/G/F",
        "nsb stage 2: forward y FFT + RK + stage 3 inverse", "nsb stage 3: chain B/G/F",
        "nsb stage 3: forward y FFT + RK + stage 4 inverse", "nsb stage 4: chain B/G/F",
        "nsb stage 4: forward y FFT + RK final combine"};
    NsbRange range(names[ph]);
    const bool device_dt = !(dt > 0.0);
    CK(cudaSetDevice(c->device));
    if (ph == 0) {
        if (sync_k(c)) return -1;
        if (c->mg_world > 1) /* peers read it during the previous step's CFL, long before this point */
            CK(cudaMemsetAsync(&c->d_sc->umax_bits, 0, sizeof(unsigned long long), c->stream));
        if (!device_dt) {
            auto k = k_set_dt;
            NSB_LAUNCH(k, dim3(1), dim3(32), 0, c->stream, c->d_sc, dt);
            CKL();
            c->launches++;
        }
        AlphaFastArgs fa;
        memset(&fa, 0, sizeof(fa));
        fill_alpha_fast(c, fa, nullptr, 0);
        fa.Wsrc = c->kW0;
        fill_rk(c, fa.rk, 1);
        return launch_alpha1024(c, false, true, fa);
    }
    if (ph & 1) { /* chain of stage (ph + 1) / 2 */
        return launch_chain(c, device_dt && ph == 1);
    }
    const int stage = ph / 2;
    if (stage == 1 && device_dt && !c->cfl_done && launch_cfl(c)) return -1;
    c->cfl_done = 0;
    AlphaFastArgs fa;
    memset(&fa, 0, sizeof(fa));
    fill_alpha_fast(c, fa, c->bufTp, c->mg_cb);
    fill_rk(c, fa.rk, stage);
    if (launch_alpha1024(c, true, stage < 4, fa)) return -1;
    if (stage == 4) {
        auto k = k_step_end;
        NSB_LAUNCH(k, dim3(1), dim3(32), 0, c->stream, c->d_sc);
        CKL();
        c->launches++;
    }
    return 0;
}

/* one process per device: every device announces `epoch` in the scalar block of every peer
 * and waits until all peers have announced it to this device (stream ordered, no host) */
static int peer_barrier_ipc(nsb_ctx *c) {
    NsbPeerScalars ps;
    for (int p = 0; p < NSB_MAXP; ++p) ps.sc[p] = c->mg_sc[p < c->mg_world ? p : 0];
    c->mg_epoch += 1;
    auto k = k_peer_barrier;
    NSB_LAUNCH(k, dim3(1), dim3(32), 0, c->stream, c->d_sc, ps, c->mg_rank, c->mg_world, c->mg_epoch);
    CKL();
    c->launches++;
    return 0;
}

static bool fast_chain_ok(const nsb_ctx *c) { return alpha_fast_ok(c) && c->chain_mode == 1; }

static int rk4_step_launch(nsb_ctx *c, double dt) {
    NsbRange range("nsb_rk4_step");
    if (fast_chain_ok(c)) {
        if (c->mg_world > 1 && !c->mg_ipc)
            return fail("this context is part of a single-process device group: use nsb_mgpu_group_step");
        for (int ph = 0; ph < NSB_FAST_PHASES; ++ph) {
            if (fast_phase(c, ph, dt)) return -1;
            if (c->mg_world > 1 && ph + 1 < NSB_FAST_PHASES && peer_barrier_ipc(c)) return -1;
        }
        return 0;
    }
    if (c->mg_world > 1)
        return fail("slab decomposition needs the fast path (Y1 == X2 == 1024, masked modes zero, NSB_CHAIN=1)");
    void *w0, *wst, *acc, *sfd = nullptr;
    if (get_field(c, NSB_OMEGA_HAT, &w0) || get_field(c, NSB_OMEGA_STAGE, &wst) || get_field(c, NSB_RK_ACC, &acc))
        return -1;
    if (c->hv.sfd && get_field(c, NSB_SFD_HAT, &sfd)) return -1;
    const bool device_dt = !(dt > 0.0);
    const bool afast = alpha_fast_ok(c);
    if (afast) {
        if (sync_k(c)) return -1;
    } else if (sync_ref(c)) {
        return -1;
    }
    if (!device_dt) {
        auto k = k_set_dt;
        NSB_LAUNCH(k, dim3(1), dim3(32), 0, c->stream, c->d_sc, dt);
        CKL();
        c->launches++;
    }
    for (int stage = 1; stage <= 4; ++stage) {
        AlphaFwdArgs af;
        memset(&af, 0, sizeof(af));
        NsbRk &rk = af.rk;
        rk.stage = stage;
        rk.W0 = (cd *)w0; rk.Wst = (cd *)wst; rk.ACC = (cd *)acc; rk.SFD = (cd *)sfd;
        rk.dt = &c->d_sc->dt;
        rk.inv_Re = 1.0 / c->prm.Re;
        rk.hv = c->hv;
        rk.FB = c->force_on ? c->d_force : nullptr;
        rk.ib = c->ib; rk.jb = c->jb;
        const cd *src = (stage == 1) ? (const cd *)w0 : (const cd *)wst;
        /* the CFL kernel must sit between gamma (stage 1) and the first RK update */
        const bool want_umax = device_dt && stage == 1;
        const NsbGeom &g = c->g;
        if (afast) {
            rk.W0 = c->kW0; rk.Wst = c->kWst; rk.ACC = c->kACC; rk.SFD = c->kSFD;
            if (stage == 1) {
                AlphaFastArgs fa;
                memset(&fa, 0, sizeof(fa));
                fill_alpha_fast(c, fa, nullptr, 0);
                fa.Wsrc = c->kW0; fa.rk = rk;
                if (launch_alpha1024(c, false, true, fa)) return -1;
            }
        } else {
            AlphaInvArgs ai;
            memset(&ai, 0, sizeof(ai));
            ai.nf = 4; ai.src[0] = src;
            ai.out = c->bufA; ai.out_fs = g.fsT; ai.out_pitch = g.PT;
            if (launch_alpha_inv(c, NSB_SRC_STEP, ai)) return -1;
        }
        if (c->chain_mode) {
            /* T (bufA) -> T' (bufB) in one persistent kernel (or three plain ones) */
            if (launch_chain(c, want_umax)) return -1;
            if (want_umax && launch_cfl(c)) return -1;
            if (afast) {
                AlphaFastArgs fa;
                memset(&fa, 0, sizeof(fa));
                fill_alpha_fast(c, fa, c->bufTp, c->mg_cb);
                fa.rk = rk;
                if (launch_alpha1024(c, true, stage < 4, fa)) return -1;
            } else {
                af.in = c->bufTp;
                af.in_pitch = c->mg_cb;
                if (launch_alpha_fwd(c, NSB_EPI_RK, af)) return -1;
                if (c->masked_nonzero) {
                    auto k = k_rk_masked;
                    NSB_LAUNCH(k, dim3(pointwise_grid(c)), dim3(256), 0, c->stream, g, rk);
                    CKL();
                    c->launches++;
                }
            }
            continue;
        }
        const cd *H = c->bufA;
        cd *Hp = c->bufB;
        if (c->has_beta) {
            BetaArgs b; b.in = c->bufA; b.out = c->bufB; b.nf = 4;
            if (launch_beta(c, true, b)) return -1;
            H = c->bufB; Hp = c->bufA;
        }
        GammaArgs ga;
        memset(&ga, 0, sizeof(ga));
        ga.in = H;
        ga.nvalid_in = (g.X1 == 1) ? g.ncol : g.X2h + 1;
        ga.out = Hp; ga.nvalid_out = ga.nvalid_in; ga.nf = 4;
        ga.umax_bits = want_umax ? &c->d_sc->umax_bits : nullptr;
        if (launch_gamma(c, NSB_GAMMA_STEP, ga)) return -1;
        if (want_umax && launch_cfl(c)) return -1;
        if (afast) {
            /* beta_fwd, then ONE kernel: forward column FFTs + RK update of this stage +
             * inverse column FFTs of the next stage */
            const cd *Tp = Hp;
            if (c->has_beta) {
                BetaArgs b; b.in = Hp; b.out = c->bufB; b.nf = 1;
                if (launch_beta(c, false, b)) return -1;
                Tp = c->bufB;
            }
            AlphaFastArgs fa;
            memset(&fa, 0, sizeof(fa));
            fill_alpha_fast(c, fa, Tp, g.PT);
            fa.rk = rk;
            if (launch_alpha1024(c, true, stage < 4, fa)) return -1;
            continue;
        }
        if (forward_tail(c, Hp, NSB_EPI_RK, af)) return -1;
        if (c->masked_nonzero) {
            auto k = k_rk_masked;
            NSB_LAUNCH(k, dim3(pointwise_grid(c)), dim3(256), 0, c->stream, g, rk);
            CKL();
            c->launches++;
        }
    }
    auto k = k_step_end;
    NSB_LAUNCH(k, dim3(1), dim3(32), 0, c->stream, c->d_sc);
    CKL();
    c->launches++;
    return 0;
}

/*
 * ns_spectral_advance_rk4 with the CFL step computed on the device is a fixed sequence of 11-55 kernel launches
 * whose arguments only change when the caller changes the configuration; on small grids (BASELINE config 2:
 * 1024^2, 0.3 ms per step) the launches themselves are a visible share of the step. The sequence is therefore
 * captured once into a CUDA graph and replayed; everything the kernels read that varies
 * from step to step (dt, t, max |u|, the forcing box) already lives in device memory.
 */
#ifndef NSB_EMULATE
static void graph_make_key(const nsb_ctx *c, unsigned char key[192]) {
    memset(key, 0, 192);
    size_t o = 0;
    auto put = [&](const void *p, size_t n) { memcpy(key + o, p, n); o += n; };
    put(&c->hv, sizeof(c->hv));
    put(&c->prm.Re, sizeof(double));
    put(&c->prm.cfl_factor, sizeof(double));
    const int ints[8] = {c->force_on, c->ib, c->jb, c->masked_nonzero, c->fast, c->chain_mode, c->mg_world, c->k_truth};
    put(ints, sizeof(ints));
    const void *ptrs[6] = {c->d_force, c->kSFD, c->kW0, c->kWst, c->kACC, c->bufP};
    put(ptrs, sizeof(ptrs));
}
#endif

extern "C" int nsb_rk4_step(nsb_ctx *c, double dt) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->device));
#ifndef NSB_EMULATE
    const bool eligible = !(dt > 0.0) && c->mg_world == 1 && !c->prof_on && c->graph_mode != 0 &&
                          (c->graph_mode > 0 || c->N <= (size_t)4096 * 4096);
    if (eligible) {
        /* the state layout the step works on must be current BEFORE the key is taken / the capture starts */
        if (alpha_fast_ok(c)) { if (sync_k(c)) return -1; } else if (sync_ref(c)) return -1;
        unsigned char key[192];
        graph_make_key(c, key);
        if (memcmp(key, c->graph_key, sizeof(key)) != 0) {
            memcpy(c->graph_key, key, sizeof(key));
            c->graph_warm = 0;
            if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
        }
        if (c->graph_exec) {
            CK(cudaGraphLaunch(c->graph_exec, c->stream));
            c->launches += c->graph_launches;
            return 0;
        }
        if (c->graph_warm >= 1) { /* one plain step has done every lazy allocation / attribute call: capture now */
            const long long l0 = c->launches;
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            const int rc = rk4_step_launch(c, dt);
            const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
            if (rc != 0 || e != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                c->launches = l0;
                c->graph_mode = 0; /* not capturable here: plain launches from now on */
                if (rc != 0) return -1;
                return rk4_step_launch(c, dt);
            }
            c->graph_launches = c->launches - l0;
            c->launches = l0;
            const cudaError_t ei = cudaGraphInstantiate(&c->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ei != cudaSuccess) {
                cudaGetLastError();
                c->graph_exec = nullptr;
                c->graph_mode = 0;
                return rk4_step_launch(c, dt);
            }
            CK(cudaGraphLaunch(c->graph_exec, c->stream));
            c->launches += c->graph_launches;
            return 0;
        }
        c->graph_warm += 1;
    }
#endif
    return rk4_step_launch(c, dt);
}

extern "C" int nsb_diagnostics(nsb_ctx *c, double out[4]) {
    if (!c || !out) return fail("nsb_diagnostics: null argument");
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    void *w;
    if (get_field(c, NSB_OMEGA_HAT, &w)) return -1;
    CK(cudaMemsetAsync(&c->d_sc->umax_bits, 0, sizeof(unsigned long long), c->stream));
    CK(cudaMemsetAsync(c->d_sc->sums, 0, sizeof(double) * 4, c->stream));
    const cd *src[4] = {(const cd *)w, nullptr, nullptr, nullptr};
    double *dst[4] = {nullptr, nullptr, nullptr, nullptr};
    if (inverse_fields(c, NSB_SRC_VEL, 3, src, dst, true, true)) return -1;
    if (fetch_scalars(c)) return -1;
    out[0] = c->h_sc->sums[0]; out[1] = c->h_sc->sums[1]; out[2] = c->h_sc->sums[2];
    double m;
    memcpy(&m, &c->h_sc->umax_bits, sizeof(double));
    out[3] = m;
    CK(cudaMemsetAsync(&c->d_sc->umax_bits, 0, sizeof(unsigned long long), c->stream));
    return 0;
}

__global__ void __launch_bounds__(256) k_sumsq3(const double *u, const double *v, const double *w, size_t n,
                                                NsbScalars *sc) {
    NSB_DYN_SMEM(double, red);
    double su = 0.0, sv = 0.0, sw = 0.0, m = 0.0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const double a = u[k], b = v[k], o = w[k];
        su += a * a; sv += b * b; sw += o * o;
        m = fmax(m, a * a + b * b);
    }
    nsb_block_sum_to_global(su, &sc->sums[0], red);
    nsb_block_sum_to_global(sv, &sc->sums[1], red);
    nsb_block_sum_to_global(sw, &sc->sums[2], red);
    nsb_block_max_to_global(m, &sc->umax_bits, red);
}

extern "C" int nsb_diagnostics_phys(nsb_ctx *c, double out[4]) {
    if (!c || !out) return fail("nsb_diagnostics_phys: null argument");
    CK(cudaSetDevice(c->device));
    void *u, *v, *w;
    if (get_field(c, NSB_U, &u) || get_field(c, NSB_V, &v) || get_field(c, NSB_OMEGA, &w)) return -1;
    CK(cudaMemsetAsync(&c->d_sc->umax_bits, 0, sizeof(unsigned long long), c->stream));
    CK(cudaMemsetAsync(c->d_sc->sums, 0, sizeof(double) * 4, c->stream));
    size_t blocks = (c->N + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    auto k = k_sumsq3;
    NSB_LAUNCH(k, dim3((unsigned)blocks), dim3(256), 256 * sizeof(double), c->stream, (const double *)u,
               (const double *)v, (const double *)w, c->N, c->d_sc);
    CKL();
    c->launches++;
    if (fetch_scalars(c)) return -1;
    out[0] = c->h_sc->sums[0]; out[1] = c->h_sc->sums[1]; out[2] = c->h_sc->sums[2];
    double m;
    memcpy(&m, &c->h_sc->umax_bits, sizeof(double));
    out[3] = m;
    CK(cudaMemsetAsync(&c->d_sc->umax_bits, 0, sizeof(unsigned long long), c->stream));
    return 0;
}

__global__ void __launch_bounds__(256) k_phys_absmax(const double *f, size_t n, unsigned long long *bits) {
    NSB_DYN_SMEM(double, red);
    double m = 0.0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        m = fmax(m, fabs(f[k]));
    nsb_block_max_to_global(m, bits, red);
}

__global__ void __launch_bounds__(256) k_phys_scale(double *f, size_t n, double s) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        f[k] = __dmul_rn(f[k], s);
}

extern "C" int nsb_phys_absmax(nsb_ctx *c, int field, double *out) {
    if (!c || !out) return fail("nsb_phys_absmax: null argument");
    if (!field_is_phys(field)) return fail("nsb_phys_absmax: not a physical field");
    CK(cudaSetDevice(c->device));
    void *f;
    if (get_field(c, field, &f)) return -1;
    CK(cudaMemsetAsync(&c->d_sc->umax_bits, 0, sizeof(unsigned long long), c->stream));
    size_t blocks = (c->N + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    auto k = k_phys_absmax;
    NSB_LAUNCH(k, dim3((unsigned)blocks), dim3(256), 256 * sizeof(double), c->stream, (const double *)f, c->N,
               &c->d_sc->umax_bits);
    CKL();
    c->launches++;
    if (fetch_scalars(c)) return -1;
    memcpy(out, &c->h_sc->umax_bits, sizeof(double));
    CK(cudaMemsetAsync(&c->d_sc->umax_bits, 0, sizeof(unsigned long long), c->stream));
    return 0;
}

extern "C" int nsb_phys_scale(nsb_ctx *c, int field, double s) {
    if (!c) return fail("null context");
    if (!field_is_phys(field)) return fail("nsb_phys_scale: not a physical field");
    CK(cudaSetDevice(c->device));
    void *f;
    if (get_field(c, field, &f)) return -1;
    size_t blocks = (c->N + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    auto k = k_phys_scale;
    NSB_LAUNCH(k, dim3((unsigned)blocks), dim3(256), 0, c->stream, (double *)f, c->N, s);
    CKL();
    c->launches++;
    return 0;
}

extern "C" int nsb_get_state(nsb_ctx *c, double *dt, double *t, int *steps) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->device));
    if (fetch_scalars(c)) return -1;
    if (dt) *dt = c->h_sc->dt;
    if (t) *t = c->h_sc->t;
    if (steps) *steps = c->h_sc->step;
    return 0;
}

extern "C" int nsb_check_finite(nsb_ctx *c, int *step_out) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->device));
    if (fetch_scalars(c)) return -1;
    if (step_out) *step_out = c->h_sc->nonfinite ? c->h_sc->nonfinite_step : -1;
    return c->h_sc->nonfinite ? 1 : 0;
}

extern "C" int nsb_set_time(nsb_ctx *c, double t) {
    if (!c) return fail("null context");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(&c->d_sc->t, &t, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int nsb_get_dt_history(nsb_ctx *c, int n, double *out) {
    if (!c || !out) return fail("nsb_get_dt_history: null argument");
    if (n < 0 || n > 256) return fail("nsb_get_dt_history: n must be 0..256");
    CK(cudaSetDevice(c->device));
    if (fetch_scalars(c)) return -1;
    const int step = c->h_sc->step;
    if (n > step) return fail("nsb_get_dt_history: only %d steps taken", step);
    for (int k = 0; k < n; ++k) out[k] = c->h_sc->dt_hist[(step - n + k) & 255];
    return 0;
}

extern "C" int nsb_get_plan(nsb_ctx *c, int out[8]) {
    if (!c || !out) return fail("nsb_get_plan: null argument");
    const NsbGeom &g = c->g;
    out[0] = g.Y1; out[1] = g.Y2; out[2] = g.X1; out[3] = g.X2;
    out[4] = g.C; out[5] = g.BC; out[6] = g.ncol;
    out[7] = 4 * (c->has_beta ? 5 : 3) + 2;
    return 0;
}

/* ------------------------------------------------------------- slab decomposition */

extern "C" int nsb_mgpu_setup(nsb_ctx *c, int rank, int world) {
    if (!c) return fail("null context");
    if (world < 1 || world > NSB_MAXP || rank < 0 || rank >= world) return fail("bad rank %d / world %d", rank, world);
    if ((world & (world - 1)) != 0) return fail("world size must be a power of two (got %d)", world);
    CK(cudaSetDevice(c->device));
    const NsbGeom &g = c->g;
    if (world > 1) {
        if (!(c->fast && g.Y1 == 1024 && c->chain_mode == 1))
            return fail("slab decomposition needs Y1 == X2 == 1024 with a beta split (grids 2048..32768) and NSB_CHAIN=1");
        if ((g.ny / world) % c->ch_rows != 0) return fail("ny / world must be a multiple of %d", c->ch_rows);
    }
    if (sync_ref(c)) return -1;
    CK(cudaStreamSynchronize(c->stream));
    if (c->kW0) { cudaFree(c->kW0); cudaFree(c->kWst); cudaFree(c->kACC); c->kW0 = c->kWst = c->kACC = nullptr; }
    if (c->kSFD) { cudaFree(c->kSFD); c->kSFD = nullptr; }
    c->k_truth = 0;
    c->mg_rank = rank;
    c->mg_world = world;
    if (world == 1) {
        c->mg_cb = g.PT; c->mg_c0 = 0; c->mg_c1 = g.ncol; c->mg_rows = g.ny; c->mg_row_shift = 30;
    } else {
        c->mg_cb = ((g.ncol + world - 1) / world + 7) / 8 * 8; /* tiles of 4 or 8 columns never straddle two devices */
        c->mg_c0 = rank * c->mg_cb;
        c->mg_c1 = c->mg_c0 + c->mg_cb < g.ncol ? c->mg_c0 + c->mg_cb : g.ncol;
        if (c->mg_c1 < c->mg_c0) c->mg_c1 = c->mg_c0;
        c->mg_rows = g.ny / world;
        int sh = 0;
        while ((1 << sh) < c->mg_rows) ++sh;
        c->mg_row_shift = sh;
    }
    for (int p = 0; p < NSB_MAXP; ++p) { c->mg_T[p] = nullptr; c->mg_Tp[p] = nullptr; c->mg_sc[p] = nullptr; }
    c->mg_T[rank] = c->bufA; c->mg_Tp[rank] = c->bufTp; c->mg_sc[rank] = c->d_sc;
    if (world == 1) { c->mg_T[0] = c->bufA; c->mg_Tp[0] = c->bufTp; c->mg_sc[0] = c->d_sc; }
    c->tmT_ok = 0;
    c->tmT8_ok = 0;
    c->tmA_ok = 0;
    c->mg_ipc = 0;
    c->mg_epoch = 0;
    /* measured on 2 B200s: 128-byte peer stores (8 columns per CTA) beat 64-byte ones by ~25 % in the
     * column kernel; on one device two 4-column CTAs per SM are slightly faster */
    if (!getenv("NSB_ALPHA_COLS")) c->alpha_cols = world > 1 ? 8 : 4;
    CK(cudaMemset(c->d_sc->peer_flags, 0, sizeof(c->d_sc->peer_flags)));
    return 0;
}

extern "C" int nsb_mgpu_partition(nsb_ctx *c, int out[6]) {
    if (!c || !out) return fail("nsb_mgpu_partition: null argument");
    out[0] = c->mg_rank; out[1] = c->mg_world; out[2] = c->mg_c0; out[3] = c->mg_c1;
    out[4] = c->mg_rank * c->mg_rows; out[5] = c->mg_rows;
    return 0;
}

/* CUDA IPC handles of the three buffers peers write to / read from: T, T', scalars */
extern "C" int nsb_mgpu_export(nsb_ctx *c, unsigned char *handles /* 3 x 64 bytes */) {
    if (!c || !handles) return fail("nsb_mgpu_export: null argument");
#ifdef NSB_EMULATE
    return fail("CUDA IPC is not available under the emulator");
#else
    CK(cudaSetDevice(c->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *ptrs[3] = {c->bufA, c->bufTp, c->d_sc};
    for (int k = 0; k < 3; ++k) {
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, ptrs[k]));
        memcpy(handles + 64 * k, &h, 64);
    }
    return 0;
#endif
}

extern "C" int nsb_mgpu_import(nsb_ctx *c, int peer, const unsigned char *handles) {
    if (!c || !handles) return fail("nsb_mgpu_import: null argument");
    if (peer < 0 || peer >= c->mg_world) return fail("bad peer %d", peer);
    if (peer == c->mg_rank) return 0;
#ifdef NSB_EMULATE
    return fail("CUDA IPC is not available under the emulator");
#else
    CK(cudaSetDevice(c->device));
    void *ptrs[3];
    for (int k = 0; k < 3; ++k) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * k, 64);
        CK(cudaIpcOpenMemHandle(&ptrs[k], h, cudaIpcMemLazyEnablePeerAccess));
        c->mg_opened[c->mg_nopened++] = ptrs[k];
    }
    c->mg_T[peer] = (cd *)ptrs[0];
    c->mg_Tp[peer] = (cd *)ptrs[1];
    c->mg_sc[peer] = (NsbScalars *)ptrs[2];
    c->mg_ipc = 1;
    c->tmA_ok = 0;
    return 0;
#endif
}

/* single process driving n devices: enable peer access and exchange raw pointers */
extern "C" int nsb_mgpu_link_local(nsb_ctx **cs, int n) {
    if (!cs || n < 1 || n > NSB_MAXP) return fail("nsb_mgpu_link_local: bad arguments");
    for (int i = 0; i < n; ++i) {
        if (!cs[i] || cs[i]->mg_world != n || cs[i]->mg_rank != i)
            return fail("nsb_mgpu_link_local: context %d is not set up as rank %d of %d", i, i, n);
    }
    for (int i = 0; i < n; ++i) {
        CK(cudaSetDevice(cs[i]->device));
        for (int j = 0; j < n; ++j) {
            if (i != j && cs[i]->device != cs[j]->device) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(cs[j]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail("cudaDeviceEnablePeerAccess(%d -> %d): %s", cs[i]->device, cs[j]->device,
                                cudaGetErrorString(e));
                cudaGetLastError();
            }
            cs[i]->mg_T[j] = cs[j]->bufA;
            cs[i]->mg_Tp[j] = cs[j]->bufTp;
            cs[i]->mg_sc[j] = cs[j]->d_sc;
        }
        cs[i]->mg_ipc = 0;
        cs[i]->tmA_ok = 0;
    }
    return 0;
}

/* all streams of the group wait for everything recorded so far on all of them */
static int group_barrier(nsb_ctx **cs, int n) {
    for (int i = 0; i < n; ++i) {
        CK(cudaSetDevice(cs[i]->device));
        CK(cudaEventRecord(cs[i]->mg_ev, cs[i]->stream));
    }
    for (int j = 0; j < n; ++j) {
        CK(cudaSetDevice(cs[j]->device));
        for (int i = 0; i < n; ++i)
            if (i != j) CK(cudaStreamWaitEvent(cs[j]->stream, cs[i]->mg_ev, 0));
    }
    return 0;
}

/* ns_spectral_advance_rk4 on a single-process device group (asynchronous) */
extern "C" int nsb_mgpu_group_step(nsb_ctx **cs, int n, double dt) {
    if (!cs || n < 1 || n > NSB_MAXP) return fail("nsb_mgpu_group_step: bad arguments");
    for (int i = 0; i < n; ++i)
        if (!cs[i] || !fast_chain_ok(cs[i]) || cs[i]->mg_world != n)
            return fail("nsb_mgpu_group_step: context %d is not on the slab-decomposed fast path", i);
    for (int ph = 0; ph < NSB_FAST_PHASES; ++ph) {
        for (int i = 0; i < n; ++i)
            if (fast_phase(cs[i], ph, dt)) return -1;
        if (n > 1 && ph + 1 < NSB_FAST_PHASES && group_barrier(cs, n)) return -1;
    }
    return 0;
}

/*
 * The step in two halves with the CFL step size handed to the host in between (8 bytes): coloured forcing needs
 * dt BEFORE the stream is consumed (/root/reference/src/common/forcing.c:138-147) but the device only knows it
 * after stage 1's physical-space kernel. begin = stage-1 transforms + max|u| + CFL kernel (phases 0-1), returns dt;
 * the caller generates the forcing, nsb_set_forcing()s it, then finish = the RK updates and stages 2-4.
 * One context (n == 1) or a single-process device group; fast path only (otherwise < 0 and nothing was launched).
 */
extern "C" int nsb_rk4_step_begin(nsb_ctx **cs, int n, double *dt_out) {
    if (!cs || n < 1 || n > NSB_MAXP || !dt_out) return fail("nsb_rk4_step_begin: bad arguments");
    for (int i = 0; i < n; ++i)
        if (!cs[i] || !fast_chain_ok(cs[i]) || cs[i]->mg_world != n || (n > 1 && cs[i]->mg_ipc))
            return fail("nsb_rk4_step_begin: context %d is not on the (slab-decomposed) fast path", i);
    for (int ph = 0; ph < 2; ++ph) {
        for (int i = 0; i < n; ++i)
            if (fast_phase(cs[i], ph, 0.0)) return -1;
        if (n > 1 && group_barrier(cs, n)) return -1;
    }
    for (int i = 0; i < n; ++i) {
        CK(cudaSetDevice(cs[i]->device));
        if (launch_cfl(cs[i])) return -1;
        cs[i]->cfl_done = 1;
    }
    CK(cudaSetDevice(cs[0]->device));
    if (fetch_scalars(cs[0])) return -1;
    *dt_out = cs[0]->h_sc->dt;
    return 0;
}

extern "C" int nsb_rk4_step_finish(nsb_ctx **cs, int n) {
    if (!cs || n < 1 || n > NSB_MAXP) return fail("nsb_rk4_step_finish: bad arguments");
    for (int i = 0; i < n; ++i)
        if (!cs[i] || !cs[i]->cfl_done) return fail("nsb_rk4_step_finish: no step in progress on context %d", i);
    for (int ph = 2; ph < NSB_FAST_PHASES; ++ph) {
        for (int i = 0; i < n; ++i)
            if (fast_phase(cs[i], ph, 0.0)) return -1;
        if (n > 1 && ph + 1 < NSB_FAST_PHASES && group_barrier(cs, n)) return -1;
    }
    return 0;
}

/* Single-process group: make cs[0]'s reference-layout OMEGA_HAT (and SFD_HAT) complete and current
 * from the own-column blocks of all devices, so that every single-device entry point
 * (diagnostics, transforms, download) can run on cs[0]. The other devices keep their own
 * columns current, so stepping can continue without a scatter. Synchronous. */
extern "C" int nsb_mgpu_group_gather(nsb_ctx **cs, int n) {
    if (!cs || n < 1 || n > NSB_MAXP) return fail("nsb_mgpu_group_gather: bad arguments");
    for (int r = 0; r < n; ++r) {
        if (!cs[r]) return fail("nsb_mgpu_group_gather: null context");
        CK(cudaSetDevice(cs[r]->device));
        if (sync_ref(cs[r])) return -1;
        CK(cudaStreamSynchronize(cs[r]->stream));
    }
    const int fields[2] = {NSB_OMEGA_HAT, NSB_SFD_HAT};
    for (int fi = 0; fi < 2; ++fi) {
        if (fi == 1 && !cs[0]->hv.sfd) break;
        void *dst0;
        CK(cudaSetDevice(cs[0]->device));
        if (get_field(cs[0], fields[fi], &dst0)) return -1;
        for (int r = 1; r < n; ++r) {
            void *src;
            CK(cudaSetDevice(cs[r]->device));
            if (get_field(cs[r], fields[fi], &src)) return -1;
            const int w = cs[r]->mg_c1 - cs[r]->mg_c0;
            if (w <= 0) continue;
            const size_t pitch = (size_t)cs[r]->g.nkx * sizeof(cd);
            CK(cudaMemcpy2DAsync((cd *)dst0 + cs[r]->mg_c0, pitch, (const cd *)src + cs[r]->mg_c0, pitch,
                                 (size_t)w * sizeof(cd), (size_t)cs[r]->g.ny, cudaMemcpyDefault, cs[r]->stream));
            CK(cudaStreamSynchronize(cs[r]->stream));
        }
    }
    return 0;
}

/* Single-process group: cs[0]'s reference-layout OMEGA_HAT (and SFD_HAT) was written by the host
 * or by a single-device entry point; copy it to every other device. Synchronous. */
extern "C" int nsb_mgpu_group_scatter(nsb_ctx **cs, int n) {
    if (!cs || n < 1 || n > NSB_MAXP) return fail("nsb_mgpu_group_scatter: bad arguments");
    CK(cudaSetDevice(cs[0]->device));
    if (sync_ref(cs[0])) return -1;
    CK(cudaStreamSynchronize(cs[0]->stream));
    const int fields[2] = {NSB_OMEGA_HAT, NSB_SFD_HAT};
    for (int r = 1; r < n; ++r) {
        CK(cudaSetDevice(cs[r]->device));
        if (sync_ref(cs[r])) return -1;
        for (int fi = 0; fi < 2; ++fi) {
            if (fi == 1 && !cs[0]->hv.sfd) break;
            void *src, *dst;
            CK(cudaSetDevice(cs[0]->device));
            if (get_field(cs[0], fields[fi], &src)) return -1;
            CK(cudaSetDevice(cs[r]->device));
            if (get_field(cs[r], fields[fi], &dst)) return -1;
            CK(cudaMemcpyAsync(dst, src, cs[0]->M * sizeof(cd), cudaMemcpyDefault, cs[r]->stream));
        }
        CK(cudaStreamSynchronize(cs[r]->stream));
        cs[r]->k_truth = 0;
        cs[r]->masked_nonzero = cs[0]->masked_nonzero;
    }
    return 0;
}

/* ---------------------------------------------------- energy / enstrophy by Parseval (device groups) */

static void reduction_columns(const nsb_ctx *c, int *ca, int *cb);

/* sum over the grid of u^2, v^2, omega^2 (ns_spectral_compute_energy / _enstrophy, spectral_analysis.c:8-33) WITHOUT
 * the three c2r transforms, from the modes of columns [ca, cb): mean(f^2) = sum_j |P f_hat(0, j)|^2 +
 * 2 sum_{i>=1} sum_j |f_hat(i, j)|^2 over the modes the 2/3 mask keeps, where P is the Hermitian projection
 * (f(j) + conj f(-j)) / 2 of the kx = 0 column -- the c2r of spectral_transforms.c:25-32 drops the imaginary part of
 * that column after the y pass, so a non-Hermitian kx = 0 column (the forcing draws one) contributes only its
 * projection. u_hat = i ky psi_hat, v_hat = -i kx psi_hat, psi_hat = -omega_hat / k^2 (spectral_core.c:17-31).
 * acc[0..2] += N * (the three means). Partial sums of the devices of a slab-decomposed group add up. */
__global__ void __launch_bounds__(256) k_parseval_sums(NsbGeom g, const cd *W, int ca, int cb, double *acc) {
    NSB_DYN_SMEM(double, red);
    const int w = cb - ca;
    const size_t total = (size_t)w * g.ny;
    double s[3] = {0.0, 0.0, 0.0};
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(k / w), i = ca + (int)(k - (size_t)j * w);
        if (i > g.imax || nsb_iabs(nsb_band(j, g.ny)) > g.ycut) continue;
        const double kxv = g.kx[i], kyv = g.ky[j];
        const double k2 = __dadd_rn(__dmul_rn(kxv, kxv), __dmul_rn(kyv, kyv));
        cd om = W[(size_t)j * g.nkx + i];
        cd uh = nsb_ik(kyv, nsb_psi_from_w(om, k2));
        cd vh = nsb_ik(kxv, nsb_psi_from_w(om, k2));
        double wt = 2.0;
        if (i == 0) {
            const int jm = (g.ny - j) % g.ny;
            const cd omm = W[(size_t)jm * g.nkx];
            const cd uhm = nsb_ik(g.ky[jm], nsb_psi_from_w(omm, k2));
            om = cmk(0.5 * (om.x + omm.x), 0.5 * (om.y - omm.y));
            uh = cmk(0.5 * (uh.x + uhm.x), 0.5 * (uh.y - uhm.y));
            vh = cmk(0.0, 0.0);
            wt = 1.0;
        }
        s[0] += wt * (uh.x * uh.x + uh.y * uh.y);
        s[1] += wt * (vh.x * vh.x + vh.y * vh.y);
        s[2] += wt * (om.x * om.x + om.y * om.y);
    }
    const double n = (double)g.nx * (double)g.ny;
    nsb_block_sum_to_global(s[0] * n, acc + 0, red);
    nsb_block_sum_to_global(s[1] * n, acc + 1, red);
    nsb_block_sum_to_global(s[2] * n, acc + 2, red);
}

extern "C" int nsb_diagnostics_spectral(nsb_ctx *c, double out[3]) {
    if (!c || !out) return fail("nsb_diagnostics_spectral: null argument");
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    void *w;
    if (get_field(c, NSB_OMEGA_HAT, &w)) return -1;
    CK(cudaMemsetAsync(c->d_sc->sums, 0, sizeof(double) * 4, c->stream));
    int ca, cb;
    reduction_columns(c, &ca, &cb);
    if (cb > c->g.ncol) cb = c->g.ncol; /* masked columns contribute nothing */
    if (cb > ca) {
        auto k = k_parseval_sums;
        NSB_LAUNCH(k, dim3(pointwise_grid(c)), dim3(256), 256 * sizeof(double), c->stream, c->g, (const cd *)w, ca, cb,
                   c->d_sc->sums);
        CKL();
        c->launches++;
    }
    if (fetch_scalars(c)) return -1;
    out[0] = c->h_sc->sums[0]; out[1] = c->h_sc->sums[1]; out[2] = c->h_sc->sums[2];
    return 0;
}

/* ---------------------------------------------------- state fingerprint */

__global__ void __launch_bounds__(256) k_state_checksum(NsbGeom g, const cd *W, int c0, int c1, unsigned long long *out) {
    const int w = c1 - c0;
    const size_t total = (size_t)w * g.ny;
    unsigned long long x = 0ull, s = 0ull;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(k / w), i = c0 + (int)(k - (size_t)j * w);
        const cd v = W[(size_t)j * g.nkx + i];
        const unsigned long long a = (unsigned long long)__double_as_longlong(v.x);
        const unsigned long long b = (unsigned long long)__double_as_longlong(v.y);
        x ^= a ^ (b << 1 | b >> 63); /* rotate the imaginary part so that re == im does not cancel */
        s += a + 3ull * b;
    }
#ifndef NSB_EMULATE
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        x ^= __shfl_xor_sync(0xffffffffu, x, off);
        s += __shfl_xor_sync(0xffffffffu, s, off);
    }
    if ((threadIdx.x & 31) != 0) return;
#endif
    atomicXor(out, x);
    atomicAdd(out + 1, s);
}

extern "C" int nsb_state_checksum(nsb_ctx *c, unsigned long long out[2]) {
    if (!c || !out) return fail("nsb_state_checksum: null argument");
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    void *w;
    if (get_field(c, NSB_OMEGA_HAT, &w)) return -1;
    unsigned long long *d = reinterpret_cast<unsigned long long *>(c->d_sc->stats);
    CK(cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), c->stream));
    if (c->mg_c1 > c->mg_c0) {
        auto k = k_state_checksum;
        NSB_LAUNCH(k, dim3(pointwise_grid(c)), dim3(256), 0, c->stream, c->g, (const cd *)w, c->mg_c0, c->mg_c1, d);
        CKL();
        c->launches++;
    }
    CK(cudaMemcpyAsync(out, d, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

/* ---------------------------------------------------- driver-loop spectral diagnostics */

/* spectral_analysis.c:36-66 (check_resolution) and :69-100 (analyze_convergence) need four
 * reductions over omega_hat; computed here so a LAZY-mode driver loop never downloads the field.
 * out = { sum |w|^2, sum |w|^2 over sqrt(k2)/sqrt(k2_last) > 0.8, max |w|,
 *         sum |w|^2 over k2 > 0.25 k2_last, number of such modes, k2_last } */
__global__ void __launch_bounds__(256) k_spectral_stats(NsbGeom g, const cd *W, int ca, int cb, double *acc,
                                                        unsigned long long *maxbits) {
    NSB_DYN_SMEM(double, red);
    const int w = cb - ca;
    const size_t total = (size_t)w * g.ny;
    const double kxl = g.kx[g.nkx - 1], kyl = g.ky[g.ny - 1];
    const double k2_last = __dadd_rn(__dmul_rn(kxl, kxl), __dmul_rn(kyl, kyl));
    const double max_k = __dsqrt_rn(k2_last);
    double tot = 0.0, high = 0.0, peak = 0.0, tail = 0.0;
    unsigned long long cnt = 0ull;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(k / w), i = ca + (int)(k - (size_t)j * w);
        const cd v = W[(size_t)j * g.nkx + i];
        const double mag = hypot(v.x, v.y);
        const double mag2 = __dmul_rn(mag, mag);
        const double kxv = g.kx[i], kyv = g.ky[j];
        const double k2 = __dadd_rn(__dmul_rn(kxv, kxv), __dmul_rn(kyv, kyv));
        tot += mag2;
        const double kn = (max_k > 0.0) ? __ddiv_rn(__dsqrt_rn(k2), max_k) : 0.0;
        if (kn > 0.8) high += mag2;
        if (mag > peak) peak = mag;
        if (k2 > 0.25 * k2_last) { tail += mag2; ++cnt; }
    }
    nsb_block_sum_to_global(tot, acc + 0, red);
    nsb_block_sum_to_global(high, acc + 1, red);
    nsb_block_sum_to_global(tail, acc + 2, red);
    nsb_block_sum_to_global((double)cnt, acc + 3, red);
    nsb_block_max_to_global(peak, maxbits, red);
    if (blockIdx.x == 0 && threadIdx.x == 0) acc[4] = k2_last;
}

/* columns a reduction over omega_hat covers on this context: everything on a single device; on a device of a
 * slab-decomposed group its own block (the last device also takes the masked columns beyond the retained set,
 * which are zero but count as modes), so that the partial results of all devices add up to the whole */
static void reduction_columns(const nsb_ctx *c, int *ca, int *cb) {
    if (c->mg_world <= 1) { *ca = 0; *cb = c->g.nkx; return; }
    *ca = c->mg_c0;
    *cb = (c->mg_rank == c->mg_world - 1) ? c->g.nkx : c->mg_c1;
    if (*cb < *ca) *cb = *ca;
}

extern "C" int nsb_spectral_stats(nsb_ctx *c, double out[6]) {
    if (!c || !out) return fail("nsb_spectral_stats: null argument");
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    void *w;
    if (get_field(c, NSB_OMEGA_HAT, &w)) return -1;
    CK(cudaMemsetAsync(c->d_sc->sums, 0, sizeof(double) * 4, c->stream));
    CK(cudaMemsetAsync(&c->d_sc->umax_bits, 0, sizeof(unsigned long long), c->stream));
    CK(cudaMemsetAsync(c->d_sc->stats, 0, sizeof(c->d_sc->stats), c->stream));
    int ca, cb;
    reduction_columns(c, &ca, &cb);
    if (cb > ca) {
        auto k = k_spectral_stats;
        NSB_LAUNCH(k, dim3(pointwise_grid(c)), dim3(256), 256 * sizeof(double), c->stream, c->g, (const cd *)w, ca, cb,
                   c->d_sc->stats, &c->d_sc->umax_bits);
        CKL();
        c->launches++;
    }
    if (fetch_scalars(c)) return -1;
    double peak;
    memcpy(&peak, &c->h_sc->umax_bits, sizeof(double));
    out[0] = c->h_sc->stats[0]; out[1] = c->h_sc->stats[1]; out[2] = peak;
    out[3] = c->h_sc->stats[2]; out[4] = c->h_sc->stats[3]; out[5] = c->h_sc->stats[4];
    CK(cudaMemsetAsync(&c->d_sc->umax_bits, 0, sizeof(unsigned long long), c->stream));
    return 0;
}

/* spectral_analysis.c:103-142 (analyze_spectrum): shell sums of 0.5 (|u_hat|^2 + |v_hat|^2) with
 * bin = (int)(sqrt(k2) + 0.5) < nbins, and the number of modes per shell, straight from omega_hat
 * (psi_hat, u_hat, v_hat per mode as in spectral_core.c:8-34; nothing is materialised or downloaded).
 * A warp covers 32 consecutive kx of one ky row, where the shell index is non-decreasing: equal
 * neighbours are combined by a segmented shuffle scan, so one atomic per (warp, shell) reaches L2. */
__global__ void __launch_bounds__(256) k_energy_spectrum(NsbGeom g, const cd *W, int ca, int cb, int nbins, double *ek,
                                                         int *cnt) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int chunks = (cb - ca + 31) >> 5;
    const size_t nwork = (size_t)g.ny * chunks;
    for (size_t wk = (size_t)blockIdx.x * wpb + (threadIdx.x >> 5); wk < nwork; wk += (size_t)gridDim.x * wpb) {
        const int j = (int)(wk / chunks), i = ca + (int)(wk - (size_t)j * chunks) * 32 + lane;
        int bin = -1;
        double e = 0.0;
        if (i < cb) {
            const double kxv = g.kx[i], kyv = g.ky[j];
            const double k2 = __dadd_rn(__dmul_rn(kxv, kxv), __dmul_rn(kyv, kyv));
            const int b = (int)(__dsqrt_rn(k2) + 0.5);
            if (b < nbins) {
                const cd p = nsb_psi_from_w(W[(size_t)j * g.nkx + i], k2);
                const cd u = nsb_ik(kyv, p), v = nsb_ik(kxv, p);
                const double a = hypot(u.x, u.y), c = hypot(v.x, v.y);
                e = 0.5 * __dadd_rn(__dmul_rn(a, a), __dmul_rn(c, c));
                bin = b;
            }
        }
        double n = bin >= 0 ? 1.0 : 0.0;
        const double bd = (double)bin;
        /* segmented inclusive scan over runs of equal `bin` */
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int src = lane >= off ? lane - off : lane;
            const double eo = __shfl_sync(0xffffffffu, e, src);
            const double no = __shfl_sync(0xffffffffu, n, src);
            const double bo = __shfl_sync(0xffffffffu, bd, src);
            if (lane >= off && bo == bd) { e += eo; n += no; }
        }
        const double bnext = __shfl_sync(0xffffffffu, bd, lane < 31 ? lane + 1 : lane);
        if (bin >= 0 && (lane == 31 || bnext != bd)) {
            atomicAdd(&ek[bin], e);
            atomicAdd(&cnt[bin], (int)n);
        }
    }
}

extern "C" int nsb_energy_spectrum(nsb_ctx *c, int nbins, double *ek_sum, int *count) {
    if (!c || !ek_sum || !count || nbins <= 0) return fail("nsb_energy_spectrum: bad argument");
    CK(cudaSetDevice(c->device));
    if (sync_ref(c)) return -1;
    void *w;
    if (get_field(c, NSB_OMEGA_HAT, &w)) return -1;
    double *d_ek = nullptr;
    const size_t bytes = (size_t)nbins * (sizeof(double) + sizeof(int));
    CK(cudaMalloc((void **)&d_ek, bytes));
    int *d_cnt = reinterpret_cast<int *>(d_ek + nbins);
    cudaError_t e = cudaMemsetAsync(d_ek, 0, bytes, c->stream);
    if (e == cudaSuccess) {
        int ca, cb;
        reduction_columns(c, &ca, &cb);
        auto k = k_energy_spectrum;
        if (cb > ca) NSB_LAUNCH(k, dim3(pointwise_grid(c)), dim3(256), 0, c->stream, c->g, (const cd *)w, ca, cb, nbins, d_ek, d_cnt);
        e = cudaGetLastError();
        c->launches++;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(ek_sum, d_ek, (size_t)nbins * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(count, d_cnt, (size_t)nbins * sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_ek);
    if (e != cudaSuccess) return fail("nsb_energy_spectrum: %s", cudaGetErrorString(e));
    return 0;
}

/* ---------------------------------------------------- turbulence statistics */

static int stats_on_device(const double *du, const double *dv, int nx, int ny, double dx, double dy, int sep_x,
                           int sep_y, double raw[10], double *corr_x_raw, double *corr_y_raw, cudaStream_t st,
                           long long *launches) {
    const size_t n = (size_t)nx * ny;
    const size_t nscr = 10 + (size_t)sep_x + (size_t)sep_y;
    double *d_scr = nullptr;
    CK(cudaMalloc((void **)&d_scr, nscr * sizeof(double)));
    cudaError_t e = cudaMemsetAsync(d_scr, 0, nscr * sizeof(double), st);
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    std::vector<double> h(nscr, 0.0);
    if (e == cudaSuccess) {
        auto k = k_stats_sums;
        NSB_LAUNCH(k, dim3((unsigned)blocks), dim3(256), 256 * sizeof(double), st, du, dv, n, d_scr);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h.data(), d_scr, 2 * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    const double inv_total = 1.0 / (double)n;
    const double mean_u = h[0] * inv_total, mean_v = h[1] * inv_total; /* statistics.c:203-204 */
    if (e == cudaSuccess) {
        auto k = k_stats_moments;
        NSB_LAUNCH(k, dim3((unsigned)blocks), dim3(256), 256 * sizeof(double), st, du, dv, nx, ny, mean_u, mean_v, dx, dy,
                   d_scr + 2);
        e = cudaGetLastError();
    }
    size_t cblocks = blocks > 148 * 2 ? 148 * 2 : blocks;
    if (e == cudaSuccess && sep_x > 0) {
        auto k = k_stats_corr;
        NSB_LAUNCH(k, dim3((unsigned)cblocks, (unsigned)sep_x), dim3(256), 256 * sizeof(double), st, du, nx, ny, mean_u, 0,
                   d_scr + 10);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && sep_y > 0) {
        auto k = k_stats_corr;
        NSB_LAUNCH(k, dim3((unsigned)cblocks, (unsigned)sep_y), dim3(256), 256 * sizeof(double), st, du, nx, ny, mean_u, 1,
                   d_scr + 10 + sep_x);
        e = cudaGetLastError();
    }
    if (launches) *launches += 2 + (sep_x > 0) + (sep_y > 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h.data() + 2, d_scr + 2, (nscr - 2) * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_scr);
    if (e != cudaSuccess) return fail("turbulence statistics: %s", cudaGetErrorString(e));
    for (int k = 0; k < 10; ++k) raw[k] = h[k];
    for (int s = 0; s < sep_x; ++s) corr_x_raw[s] = h[10 + s];
    for (int s = 0; s < sep_y; ++s) corr_y_raw[s] = h[10 + sep_x + s];
    return 0;
}

extern "C" int nsb_velocity_statistics(nsb_ctx *c, double dx, double dy, int sep_x, int sep_y, double raw[10],
                                       double *corr_x_raw, double *corr_y_raw) {
    if (!c || !raw) return fail("nsb_velocity_statistics: null argument");
    if (sep_x < 0 || sep_y < 0 || sep_x > c->g.nx || sep_y > c->g.ny || (sep_x > 0 && !corr_x_raw) || (sep_y > 0 && !corr_y_raw))
        return fail("nsb_velocity_statistics: bad separations");
    CK(cudaSetDevice(c->device));
    void *u, *v;
    if (get_field(c, NSB_U, &u) || get_field(c, NSB_V, &v)) return -1;
    return stats_on_device((const double *)u, (const double *)v, c->g.nx, c->g.ny, dx, dy, sep_x, sep_y, raw, corr_x_raw,
                           corr_y_raw, c->stream, &c->launches);
}

extern "C" int nsb_velocity_statistics_host(int device, const double *u, const double *v, int nx, int ny, double dx,
                                            double dy, int sep_x, int sep_y, double raw[10], double *corr_x_raw,
                                            double *corr_y_raw) {
    if (!u || !v || !raw || nx <= 0 || ny <= 0) return fail("nsb_velocity_statistics_host: bad argument");
    if (sep_x < 0 || sep_y < 0 || sep_x > nx || sep_y > ny || (sep_x > 0 && !corr_x_raw) || (sep_y > 0 && !corr_y_raw))
        return fail("nsb_velocity_statistics_host: bad separations");
#ifndef NSB_EMULATE
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail("no CUDA device visible: this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail("device %d out of range (%d visible)", device, ndev);
#endif
    CK(cudaSetDevice(device));
    const size_t bytes = (size_t)nx * ny * sizeof(double);
    double *du = nullptr, *dv = nullptr;
    CK(cudaMalloc((void **)&du, bytes));
    if (cudaMalloc((void **)&dv, bytes) != cudaSuccess) { cudaFree(du); return fail("out of device memory"); }
    int rc = -1;
    if (cudaMemcpy(du, u, bytes, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(dv, v, bytes, cudaMemcpyHostToDevice) == cudaSuccess)
        rc = stats_on_device(du, dv, nx, ny, dx, dy, sep_x, sep_y, raw, corr_x_raw, corr_y_raw, (cudaStream_t)0, nullptr);
    else
        fail("nsb_velocity_statistics_host: upload failed");
    cudaFree(du);
    cudaFree(dv);
    return rc;
}
