/*
 * spectral_api.c -- the spectral solver's C entry points (include/ns_spectral_solver.h)
 * implemented over the kernel C-ABI (include/nsb.h). Host C, no numerics of its own
 * beyond setup tables and driver diagnostics: every transform, operator, reduction and
 * RK stage runs in the sm_100a kernels. No CPU fallback.
 *
 * Each function cites the reference function it replaces
 * (/root/reference/src/spectral/...). See the header for the EAGER / LAZY coherence
 * policy.
 */
#include "ns_spectral_solver.h"
#include "nsb.h"

#include <stdarg.h>

typedef struct {
    nsb_ctx *ctx;
    int device;
    NSSyncPolicy policy;
    /* ns_spectral_backend_init: slab decomposition over several devices driven by this process.
     * group[0] == ctx. `gathered`: ctx's reference-layout omega_hat is complete (single-device entry
     * points may run on it); `dirty`: it may have been changed since the other devices last saw it. */
    nsb_ctx *group[8];
    int ngpus, gathered, dirty;
    /* forcing fast path: ascending active-index list of the last NSForcingState seen */
    const NSForcingState *forcing_owner;
    int *forcing_idx;
    int forcing_n, ib, jb;
    double _Complex *forcing_box;
} Backend;

static _Thread_local char g_error[512] = "";

static void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}

const char *ns_spectral_last_error(void) { return g_error; }
/* for the other host sources of this library (analysis.c) */
void ns_spectral_set_error_text(const char *text) { set_error("%s", text ? text : ""); }

static Backend *backend_of(const NSSpectralData *d) { return d ? (Backend *)d->backend : NULL; }

/* loud failure: the reference's entry points return void, so an unusable device cannot
 * be reported through them -- and must never degrade into a silent CPU path. */
static void die(const char *where, const char *why) {
    fprintf(stderr, "ns_spectral (B200 backend): %s: %s\n", where, why);
    abort();
}

#define NSB_OK(call, where)                               \
    do {                                                  \
        if ((call) != 0) die(where, nsb_last_error());    \
    } while (0)

/* the primary context for a single-device entry point; in a device group the own-column blocks
 * of all devices are gathered into it first. `modifies`: the call may change omega_hat there. */
static nsb_ctx *ctx_get(NSSpectralData *d, const char *where, int modifies) {
    Backend *b = backend_of(d);
    if (!d || !b) die(where, "NULL solver data");
    if (!b->ctx) die(where, "no device context: call ns_spectral_initialize_fft_plans() on a machine with a "
                            "CUDA device first (there is no CPU fallback)");
    if (b->ngpus > 1) {
        if (!b->gathered) {
            NSB_OK(nsb_mgpu_group_gather(b->group, b->ngpus), where);
            b->gathered = 1;
        }
        if (modifies) b->dirty = 1;
    }
    return b->ctx;
}
static nsb_ctx *ctx_of(NSSpectralData *d, const char *where) { return ctx_get(d, where, 1); }
static nsb_ctx *ctx_ro(NSSpectralData *d, const char *where) { return ctx_get(d, where, 0); }

static size_t n_phys(const NSSpectralData *d) { return (size_t)d->nx * (size_t)d->ny; }
static size_t n_spec(const NSSpectralData *d) { return (size_t)d->nkx * (size_t)d->nky; }
static int lazy(const NSSpectralData *d) { return backend_of(d)->policy == NS_SYNC_LAZY; }

/* ------------------------------------------------------------------ memory */

/* calloc: large arrays come from fresh zero pages and cost nothing until touched
 * (the struct carries 138 bytes per grid point, 37 GB at 16384^2) */
static void *zalloc64(size_t bytes) { return calloc(bytes ? bytes : 16, 1); }

/* spectral_memory.c:6-68. All arrays zero-filled (the reference leaves them
 * uninitialised). */
NSSpectralData *ns_spectral_allocate(int nx, int ny, double Lx, double Ly, double Re) {
    if (nx <= 0 || ny <= 0) return NULL;
    NSSpectralData *d = (NSSpectralData *)calloc(1, sizeof *d);
    if (!d) return NULL;
    d->nx = nx; d->ny = ny; d->nkx = nx / 2 + 1; d->nky = ny;
    d->Lx = Lx; d->Ly = Ly; d->Re = Re;
    d->dt = 0.01; d->t = 0.0; d->cfl_factor = 0.3;
    const size_t n = n_phys(d), m = n_spec(d);
    double **re[] = {&d->u, &d->v, &d->p, &d->omega, &d->psi};
    fftw_complex **cx[] = {&d->u_hat, &d->v_hat, &d->omega_hat, &d->psi_hat, &d->nonlinear_hat, &d->rk1_hat,
                           &d->rk2_hat, &d->rk3_hat, &d->rk4_hat, &d->omega_tmp_hat, &d->sfd_hat};
    int ok = 1;
    for (size_t k = 0; k < sizeof re / sizeof re[0]; ++k) ok &= ((*re[k] = (double *)zalloc64(n * sizeof(double))) != NULL);
    for (size_t k = 0; k < sizeof cx / sizeof cx[0]; ++k) ok &= ((*cx[k] = (fftw_complex *)zalloc64(m * sizeof(fftw_complex))) != NULL);
    d->x = (double *)calloc((size_t)nx, sizeof(double));
    d->y = (double *)calloc((size_t)ny, sizeof(double));
    d->kx = (double *)calloc((size_t)d->nkx, sizeof(double));
    d->ky = (double *)calloc((size_t)d->nky, sizeof(double));
    d->k2 = (double *)calloc(m, sizeof(double));
    d->dealias_mask = (int *)calloc(m, sizeof(int));
    d->hyperviscosity_weights = (double *)calloc(m, sizeof(double));
    Backend *b = (Backend *)calloc(1, sizeof *b);
    d->backend = b;
    if (!ok || !d->x || !d->y || !d->kx || !d->ky || !d->k2 || !d->dealias_mask || !d->hyperviscosity_weights || !b) {
        ns_spectral_free(d);
        return NULL;
    }
    const char *env = getenv("NS_SPECTRAL_DEVICE");
    b->device = env ? atoi(env) : 0;
    b->policy = NS_SYNC_EAGER;
    ns_spectral_hyperviscosity_default(&d->hyperviscosity);
    return d;
}

/* spectral_memory.c:71-104 (no global fftw_cleanup: instances are independent here) */
void ns_spectral_free(NSSpectralData *d) {
    if (!d) return;
    Backend *b = backend_of(d);
    if (b) {
        for (int r = 1; r < b->ngpus; ++r)
            if (b->group[r]) nsb_ctx_destroy(b->group[r]);
        if (b->ctx) nsb_ctx_destroy(b->ctx);
        free(b->forcing_idx);
        free(b->forcing_box);
        free(b);
    }
    free(d->u); free(d->v); free(d->p); free(d->omega); free(d->psi);
    free(d->u_hat); free(d->v_hat); free(d->omega_hat); free(d->psi_hat); free(d->nonlinear_hat);
    free(d->rk1_hat); free(d->rk2_hat); free(d->rk3_hat); free(d->rk4_hat); free(d->omega_tmp_hat);
    free(d->sfd_hat);
    free(d->x); free(d->y); free(d->kx); free(d->ky); free(d->k2);
    free(d->dealias_mask); free(d->hyperviscosity_weights);
    free(d);
}

/* ---------------------------------------------------------- initialisation */

/* spectral_initialization.c:10-54 (host tables; the kernels rebuild kx, ky, k2 and the
 * mask from indices with the same formulas) */
void ns_spectral_initialize_grids(NSSpectralData *d) {
    if (!d) return;
    const int nx = d->nx, ny = d->ny, nkx = d->nkx, nky = d->nky;
    for (int i = 0; i < nx; ++i) d->x[i] = i * d->Lx / nx;
    for (int j = 0; j < ny; ++j) d->y[j] = j * d->Ly / ny;
    for (int i = 0; i < nkx; ++i) d->kx[i] = 2.0 * M_PI * i / d->Lx;
    for (int j = 0; j < nky; ++j)
        d->ky[j] = (j <= ny / 2) ? 2.0 * M_PI * j / d->Ly : 2.0 * M_PI * (j - ny) / d->Ly;
    const int xcut = (2 * nkx) / 3, ycut = (2 * ny) / 6;
    for (int j = 0; j < nky; ++j) {
        const int band = (j <= ny / 2) ? j : j - ny;
        for (int i = 0; i < nkx; ++i) {
            const size_t idx = (size_t)j * nkx + i;
            d->k2[idx] = d->kx[i] * d->kx[i] + d->ky[j] * d->ky[j];
            d->dealias_mask[idx] = (i <= xcut) && (abs(band) <= ycut);
        }
    }
    ns_spectral_hyperviscosity_update_weights(d);
}

static void push_hyperviscosity(NSSpectralData *d) {
    Backend *b = backend_of(d);
    if (!b || !b->ctx) return;
    const NSSpectralHyperviscosityConfig *h = &d->hyperviscosity;
    nsb_hyperviscosity p;
    p.enabled = (h->enabled && h->coefficient > 0.0);
    p.coefficient = h->coefficient;
    p.order = h->order > 0 ? h->order : 1;
    p.cutoff_wavenumber = h->cutoff_wavenumber;
    p.inv_bandwidth = h->inv_bandwidth;
    p.sfd_enabled = (h->sfd_enabled && h->sfd_strength > 0.0);
    p.sfd_strength = h->sfd_strength;
    p.sfd_timescale = h->sfd_timescale;
    NSB_OK(nsb_set_hyperviscosity(b->ctx, &p), "hyperviscosity");
    for (int r = 1; r < b->ngpus; ++r) NSB_OK(nsb_set_hyperviscosity(b->group[r], &p), "hyperviscosity");
}

/* the reference reads data->Re and data->cfl_factor at every use (spectral_time_integration.c:22-24,
 * 52): hand the caller's current values to every device context before a step / CFL evaluation */
static void push_physics(NSSpectralData *d, const char *where) {
    Backend *b = backend_of(d);
    if (!b || !b->ctx) return;
    NSB_OK(nsb_set_physics(b->ctx, d->Re, d->cfl_factor), where);
    for (int r = 1; r < b->ngpus; ++r) NSB_OK(nsb_set_physics(b->group[r], d->Re, d->cfl_factor), where);
}

/* spectral_initialization.c:57-69: "plan creation" = device context creation */
void ns_spectral_initialize_fft_plans(NSSpectralData *d) {
    Backend *b = backend_of(d);
    if (!d || !b) return;
    if (b->ctx) return;
    nsb_params p;
    memset(&p, 0, sizeof p);
    p.nx = d->nx; p.ny = d->ny; p.Lx = d->Lx; p.Ly = d->Ly; p.Re = d->Re;
    p.cfl_factor = d->cfl_factor;
    p.device = b->device;
    if (nsb_ctx_create(&p, &b->ctx) != 0) {
        set_error("%s", nsb_last_error());
        fprintf(stderr, "ns_spectral (B200 backend): cannot create device context: %s\n", nsb_last_error());
        b->ctx = NULL;
        d->forward_plan = d->backward_plan = NULL;
        return;
    }
    d->forward_plan = d->backward_plan = (fftw_plan)b->ctx;
    push_hyperviscosity(d);
}

static void upload(NSSpectralData *d, int field, const void *host, const char *where) {
    NSB_OK(nsb_upload(ctx_of(d, where), field, host), where);
}
static void download(NSSpectralData *d, int field, void *host, const char *where) {
    NSB_OK(nsb_download(ctx_ro(d, where), field, host), where);
}

/* device: psi_hat,u_hat,v_hat,u,v,omega from omega_hat; host arrays refreshed */
static void refresh_derived_to_host(NSSpectralData *d, const char *where) {
    nsb_ctx *c = ctx_of(d, where);
    NSB_OK(nsb_compute_derivatives(c), where);
    download(d, NSB_PSI_HAT, d->psi_hat, where);
    download(d, NSB_U_HAT, d->u_hat, where);
    download(d, NSB_V_HAT, d->v_hat, where);
    download(d, NSB_U, d->u, where);
    download(d, NSB_V, d->v, where);
    download(d, NSB_OMEGA, d->omega, where);
}

/* spectral_initialization.c:72-95 (the inconsistent omega of G8 is kept: parity) */
void ns_spectral_initialize_taylor_green(NSSpectralData *d) {
    static const char *W = "ns_spectral_initialize_taylor_green";
    nsb_ctx *c = ctx_of(d, W);
    const int nx = d->nx, ny = d->ny;
    const double Lx = d->Lx, Ly = d->Ly;
    for (int j = 0; j < ny; ++j) {
        for (int i = 0; i < nx; ++i) {
            const size_t idx = (size_t)j * nx + i;
            const double x = d->x[i], y = d->y[j];
            d->u[idx] = sin(2 * M_PI * x / Lx) * cos(2 * M_PI * y / Ly);
            d->v[idx] = -cos(2 * M_PI * x / Lx) * sin(2 * M_PI * y / Ly);
            d->omega[idx] = 2.0 * (2 * M_PI / Lx) * cos(2 * M_PI * x / Lx) * cos(2 * M_PI * y / Ly);
            d->p[idx] = 0.0;
        }
    }
    upload(d, NSB_OMEGA, d->omega, W);
    upload(d, NSB_U, d->u, W);
    upload(d, NSB_V, d->v, W);
    NSB_OK(nsb_physical_to_spectral(c, NSB_OMEGA, NSB_OMEGA_HAT), W);
    NSB_OK(nsb_physical_to_spectral(c, NSB_U, NSB_U_HAT), W);
    NSB_OK(nsb_physical_to_spectral(c, NSB_V, NSB_V_HAT), W);
    download(d, NSB_OMEGA_HAT, d->omega_hat, W);
    download(d, NSB_U_HAT, d->u_hat, W);
    download(d, NSB_V_HAT, d->v_hat, W);
}

static unsigned long long sm64(unsigned long long *st) {
    unsigned long long z = (*st += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* SURVEY 8d (build-defined; the reference has no turbulence initial condition): for idx = j*nkx + i
 * ascending over ALL modes one splitmix64 draw -> phase, omega_hat = a(k) e^{i theta}, a = k^2 exp(-(k/k0)^2);
 * c2r; scale to max|omega| = max_omega; r2c (normalise + mask + Hermitian projection); derivatives.
 * Same definition as nsref_random_phase_init of the oracle build. Where exp() underflows (a == 0
 * exactly) the trigonometry is skipped but the draw is still consumed. Under NS_SYNC_LAZY the peak
 * search and the rescaling run on the device and no field is copied back (16384^2: 2 GB up, 0 down). */
int ns_spectral_initialize_random_phase(NSSpectralData *d, unsigned long long seed, double k0, double max_omega) {
    static const char *W = "ns_spectral_initialize_random_phase";
    nsb_ctx *c = ctx_of(d, W);
    const int nkx = d->nkx, nky = d->nky, ny = d->ny;
    unsigned long long st = seed;
    for (int j = 0; j < nky; ++j) {
        const int band = (j <= ny / 2) ? j : j - ny;
        for (int i = 0; i < nkx; ++i) {
            const double un = (double)(sm64(&st) >> 11) * (1.0 / 9007199254740992.0);
            const double kk = sqrt((double)i * i + (double)band * band);
            const double e = (kk / k0) * (kk / k0);
            const size_t idx = (size_t)j * nkx + i;
            if (e > 745.2) { /* exp(-e) == 0 in binary64 */
                if (creal(d->omega_hat[idx]) != 0.0 || cimag(d->omega_hat[idx]) != 0.0) d->omega_hat[idx] = 0.0;
                continue;
            }
            const double theta = 2.0 * M_PI * un;
            const double a = kk * kk * exp(-e);
            d->omega_hat[idx] = a * cos(theta) + I * (a * sin(theta));
        }
    }
    d->omega_hat[0] = 0.0;
    upload(d, NSB_OMEGA_HAT, d->omega_hat, W);
    const int src = NSB_OMEGA_HAT, dst = NSB_OMEGA;
    NSB_OK(nsb_spectral_to_physical(c, 1, &src, &dst), W);
    double peak = 0.0;
    NSB_OK(nsb_phys_absmax(c, NSB_OMEGA, &peak), W);
    const double scale = (peak > 0.0) ? max_omega / peak : 0.0;
    NSB_OK(nsb_phys_scale(c, NSB_OMEGA, scale), W);
    NSB_OK(nsb_physical_to_spectral(c, NSB_OMEGA, NSB_OMEGA_HAT), W);
    if (lazy(d)) {
        NSB_OK(nsb_compute_derivatives(c), W);
        return 0;
    }
    download(d, NSB_OMEGA_HAT, d->omega_hat, W);
    refresh_derived_to_host(d, W);
    return 0;
}

/* -------------------------------------------------------------- transforms */

/* spectral_transforms.c:9-22 */
void ns_spectral_physical_to_spectral(NSSpectralData *d, double *field, fftw_complex *field_hat) {
    static const char *W = "ns_spectral_physical_to_spectral";
    nsb_ctx *c = ctx_of(d, W);
    if (!field || !field_hat) return;
    upload(d, NSB_SCRATCH_PHYS, field, W);
    const int dst = (field_hat == d->omega_hat) ? NSB_OMEGA_HAT : NSB_SCRATCH_HAT;
    NSB_OK(nsb_physical_to_spectral(c, NSB_SCRATCH_PHYS, dst), W);
    download(d, dst, field_hat, W);
}

/* spectral_transforms.c:25-32 (the caller's field_hat is left untouched; the
 * reference's use of nonlinear_hat as scratch is not reproduced) */
void ns_spectral_spectral_to_physical(NSSpectralData *d, fftw_complex *field_hat, double *field) {
    static const char *W = "ns_spectral_spectral_to_physical";
    nsb_ctx *c = ctx_of(d, W);
    if (!field_hat || !field) return;
    upload(d, NSB_SCRATCH_HAT, field_hat, W);
    const int src = NSB_SCRATCH_HAT, dst = NSB_SCRATCH_PHYS;
    NSB_OK(nsb_spectral_to_physical(c, 1, &src, &dst), W);
    download(d, NSB_SCRATCH_PHYS, field, W);
}

/* spectral_transforms.c:35-43 */
void ns_spectral_apply_dealiasing(NSSpectralData *d, fftw_complex *field_hat) {
    static const char *W = "ns_spectral_apply_dealiasing";
    nsb_ctx *c = ctx_of(d, W);
    if (!field_hat) return;
    if (lazy(d) && field_hat == d->omega_hat) {
        NSB_OK(nsb_apply_dealiasing(c, NSB_OMEGA_HAT), W);
        return;
    }
    upload(d, NSB_SCRATCH_HAT, field_hat, W);
    NSB_OK(nsb_apply_dealiasing(c, NSB_SCRATCH_HAT), W);
    download(d, NSB_SCRATCH_HAT, field_hat, W);
}

/* -------------------------------------------------------------------- core */

/* spectral_core.c:8-34 */
void ns_spectral_compute_velocity_from_vorticity(NSSpectralData *d) {
    static const char *W = "ns_spectral_compute_velocity_from_vorticity";
    nsb_ctx *c = ctx_of(d, W);
    if (!lazy(d)) upload(d, NSB_OMEGA_HAT, d->omega_hat, W);
    NSB_OK(nsb_velocity_from_vorticity(c), W);
    if (!lazy(d)) {
        download(d, NSB_PSI_HAT, d->psi_hat, W);
        download(d, NSB_U_HAT, d->u_hat, W);
        download(d, NSB_V_HAT, d->v_hat, W);
    }
}

/* spectral_core.c:37-43 */
void ns_spectral_compute_derivatives(NSSpectralData *d) {
    static const char *W = "ns_spectral_compute_derivatives";
    nsb_ctx *c = ctx_of(d, W);
    if (lazy(d)) {
        NSB_OK(nsb_compute_derivatives(c), W);
        return;
    }
    upload(d, NSB_OMEGA_HAT, d->omega_hat, W);
    refresh_derived_to_host(d, W);
}

/* spectral_core.c:46-112: nonlinear_hat from the STORED psi_hat and omega_hat */
void ns_spectral_compute_nonlinear_terms(NSSpectralData *d) {
    static const char *W = "ns_spectral_compute_nonlinear_terms";
    nsb_ctx *c = ctx_of(d, W);
    if (!lazy(d)) {
        upload(d, NSB_PSI_HAT, d->psi_hat, W);
        upload(d, NSB_OMEGA_HAT, d->omega_hat, W);
    }
    NSB_OK(nsb_nonlinear_terms(c), W);
    if (!lazy(d)) download(d, NSB_NONLINEAR_HAT, d->nonlinear_hat, W);
}

/* --------------------------------------------------------- time integration */

/* spectral_time_integration.c:9-31 */
double ns_spectral_compute_cfl_timestep(NSSpectralData *d) {
    static const char *W = "ns_spectral_compute_cfl_timestep";
    nsb_ctx *c = ctx_of(d, W);
    push_physics(d, W);
    double dt = 0.0;
    if (lazy(d)) {
        NSB_OK(nsb_cfl_timestep(c, &dt), W);
    } else {
        upload(d, NSB_U, d->u, W);
        upload(d, NSB_V, d->v, W);
        NSB_OK(nsb_cfl_timestep_uv(c, &dt), W);
    }
    return dt;
}

static int push_sfd_if_needed(NSSpectralData *d, const char *where) {
    const NSSpectralHyperviscosityConfig *h = &d->hyperviscosity;
    if (h->sfd_enabled && h->sfd_strength > 0.0) {
        upload(d, NSB_SFD_HAT, d->sfd_hat, where);
        return 1;
    }
    return 0;
}

/* the same forcing box on every device of a group (each applies it to its own columns) */
static void set_forcing_all(NSSpectralData *d, int ib, int jb, const double *box, const char *where) {
    Backend *b = backend_of(d);
    NSB_OK(nsb_set_forcing(b->ctx, ib, jb, box), where);
    for (int r = 1; r < b->ngpus; ++r) NSB_OK(nsb_set_forcing(b->group[r], ib, jb, box), where);
}

/* dense forcing array -> bounding box of its non-zero entries -> device(s) */
static void push_dense_forcing(NSSpectralData *d, const fftw_complex *f, const char *where) {
    if (!f) {
        set_forcing_all(d, 0, 0, NULL, where);
        return;
    }
    int ib = 0, jb = 0;
    for (int j = 0; j < d->nky; ++j) {
        const int band = (j <= d->ny / 2) ? j : j - d->ny;
        for (int i = 0; i < d->nkx; ++i) {
            const fftw_complex v = f[(size_t)j * d->nkx + i];
            if (creal(v) != 0.0 || cimag(v) != 0.0) {
                if (i > ib) ib = i;
                if (abs(band) > jb) jb = abs(band);
            }
        }
    }
    if (jb >= d->ny / 2) jb = d->ny / 2 - 1;
    const size_t cells = (size_t)(2 * jb + 1) * (size_t)(ib + 1);
    double _Complex *box = (double _Complex *)calloc(cells, sizeof *box);
    if (!box) die(where, "out of memory");
    for (int band = -jb; band <= jb; ++band) {
        const int j = (band >= 0) ? band : band + d->ny;
        for (int i = 0; i <= ib; ++i) box[(size_t)(band + jb) * (ib + 1) + i] = f[(size_t)j * d->nkx + i];
    }
    set_forcing_all(d, ib, jb, (const double *)box, where);
    free(box);
}

/* one RK4 step with step size dt (dt <= 0: CFL step computed on the device) on the single device or,
 * after ns_spectral_backend_init(n > 1), slab-decomposed on the device group: what the single-device
 * entry points did to omega_hat on the first device since the last group step is published first */
static void step_devices(NSSpectralData *d, double dt, const char *where) {
    Backend *b = backend_of(d);
    if (b->ngpus > 1) {
        if (b->dirty) NSB_OK(nsb_mgpu_group_scatter(b->group, b->ngpus), where);
        b->dirty = 0;
        NSB_OK(nsb_mgpu_group_step(b->group, b->ngpus, dt), where);
        b->gathered = 0;
    } else {
        NSB_OK(nsb_rk4_step(b->ctx, dt), where);
    }
}

static void set_time_all(NSSpectralData *d, const char *where) {
    Backend *b = backend_of(d);
    NSB_OK(nsb_set_time(b->ctx, d->t), where);
    for (int r = 1; r < b->ngpus; ++r) NSB_OK(nsb_set_time(b->group[r], d->t), where);
}

/* spectral_time_integration.c:34-106 (+ f_hat) */
int ns_spectral_advance_rk4_forced(NSSpectralData *d, const fftw_complex *forcing) {
    static const char *W = "ns_spectral_advance_rk4";
    ctx_of(d, W);
    if (!(d->dt > 0.0)) {
        set_error("advance_rk4: data->dt must be positive (got %g)", d->dt);
        return -1;
    }
    int sfd = 0;
    if (!lazy(d)) {
        upload(d, NSB_OMEGA_HAT, d->omega_hat, W);
        sfd = push_sfd_if_needed(d, W);
    }
    push_physics(d, W);
    push_dense_forcing(d, forcing, W);
    step_devices(d, d->dt, W);
    d->t += d->dt;
    if (!lazy(d)) {
        download(d, NSB_OMEGA_HAT, d->omega_hat, W);
        if (sfd) download(d, NSB_SFD_HAT, d->sfd_hat, W);
        refresh_derived_to_host(d, W);
    }
    set_time_all(d, W);
    return 0;
}

void ns_spectral_advance_rk4(NSSpectralData *d) {
    if (ns_spectral_advance_rk4_forced(d, NULL) != 0) die("ns_spectral_advance_rk4", g_error);
}

static int bind_forcing(NSSpectralData *d, NSForcingState *fs) {
    Backend *b = backend_of(d);
    if (b->forcing_owner == fs && b->forcing_idx) return 0;
    free(b->forcing_idx);
    free(b->forcing_box);
    b->forcing_idx = NULL;
    b->forcing_box = NULL;
    b->forcing_owner = NULL;
    if (fs->nkx != d->nkx || fs->nky != d->nky || fs->active_modes <= 0) {
        set_error("forcing state does not match the solver grid");
        return -1;
    }
    b->forcing_idx = (int *)malloc((size_t)fs->active_modes * sizeof(int));
    if (!b->forcing_idx || ns_forcing_active_indices(fs, b->forcing_idx) != 0 ||
        ns_forcing_box_extent(fs, &b->ib, &b->jb) != 0) {
        set_error("cannot index the forcing band");
        return -1;
    }
    if (b->jb >= d->ny / 2) b->jb = d->ny / 2 - 1;
    b->forcing_box = (double _Complex *)malloc((size_t)(2 * b->jb + 1) * (size_t)(b->ib + 1) * sizeof(double _Complex));
    if (!b->forcing_box) return -1;
    b->forcing_n = fs->active_modes;
    b->forcing_owner = fs;
    return 0;
}

/* main_spectral.c:46-48 n times, device resident */
int ns_spectral_run_steps(NSSpectralData *d, int n, NSForcingState *fs) {
    static const char *W = "ns_spectral_run_steps";
    nsb_ctx *c = ctx_of(d, W);
    Backend *b = backend_of(d);
    if (n < 0) return -1;
    int sfd = 0;
    if (!lazy(d)) {
        upload(d, NSB_OMEGA_HAT, d->omega_hat, W);
        sfd = push_sfd_if_needed(d, W);
    }
    push_physics(d, W);
    set_time_all(d, W);
    if (fs && bind_forcing(d, fs) != 0) return -1;
    const int group = b->ngpus > 1;
    if (!fs) set_forcing_all(d, 0, 0, NULL, W);
    for (int s = 0; s < n; ++s) {
        double dt = 0.0; /* 0 = CFL step computed on the device inside the step */
        if (fs) {
            /* coloured noise needs dt before the stream is consumed (forcing.c:138-147): the CFL step is
             * evaluated first (on a group: from the state gathered into the first device, 8 bytes back) and
             * handed to the step, exactly the reference's `data->dt = cfl(); advance_rk4()` order */
            if (fs->config.type == NS_FORCING_COLORED_NOISE) {
                /* fast path: the step is split around the device's own CFL kernel (8 bytes back), on one device or
                 * on a group; other grids evaluate the CFL step in a separate pass over omega_hat first */
                nsb_ctx **cs = group ? b->group : &b->ctx;
                if (group && b->dirty) {
                    NSB_OK(nsb_mgpu_group_scatter(b->group, b->ngpus), W);
                    b->dirty = 0;
                }
                if (nsb_rk4_step_begin(cs, group ? b->ngpus : 1, &dt) == 0) {
                    ns_forcing_generate_box(fs, dt, b->forcing_idx, b->ib, b->jb, b->forcing_box);
                    set_forcing_all(d, b->ib, b->jb, (const double *)b->forcing_box, W);
                    NSB_OK(nsb_rk4_step_finish(cs, group ? b->ngpus : 1), W);
                    if (group) b->gathered = 0;
                    continue;
                }
                c = ctx_ro(d, W);
                NSB_OK(nsb_cfl_timestep(c, &dt), W);
            }
            ns_forcing_generate_box(fs, dt, b->forcing_idx, b->ib, b->jb, b->forcing_box);
            set_forcing_all(d, b->ib, b->jb, (const double *)b->forcing_box, W);
        }
        step_devices(d, dt, W);
    }
    (void)group;
    int steps = 0;
    NSB_OK(nsb_get_state(b->ctx, &d->dt, &d->t, &steps), W);
    /* a blown-up field is reported, not integrated on with dt = MAX_DT */
    int bad_step = -1;
    for (int r = 0; r < (b->ngpus > 1 ? b->ngpus : 1); ++r) {
        const int bad = nsb_check_finite(b->ngpus > 1 ? b->group[r] : b->ctx, &bad_step);
        if (bad < 0) die(W, nsb_last_error());
        if (bad) {
            set_error("%s: the velocity field is no longer finite (NaN/Inf seen by the CFL reduction at step %d)", W,
                      bad_step);
            return -1;
        }
    }
    if (!lazy(d)) {
        download(d, NSB_OMEGA_HAT, d->omega_hat, W);
        if (sfd) download(d, NSB_SFD_HAT, d->sfd_hat, W);
        refresh_derived_to_host(d, W);
    }
    return 0;
}

/* ------------------------------------------------------------- diagnostics */

static int group_partial(const NSSpectralData *d);

static void sums(NSSpectralData *d, double out[4], const char *where) {
    if (group_partial(d)) { /* Parseval partial sums over every device's own columns, no gather */
        Backend *b = backend_of(d);
        out[0] = out[1] = out[2] = out[3] = 0.0;
        for (int r = 0; r < b->ngpus; ++r) {
            double part[3];
            NSB_OK(nsb_diagnostics_spectral(b->group[r], part), where);
            out[0] += part[0]; out[1] += part[1]; out[2] += part[2];
        }
        return;
    }
    nsb_ctx *c = ctx_ro(d, where);
    if (lazy(d)) {
        NSB_OK(nsb_diagnostics(c, out), where);
    } else {
        upload(d, NSB_U, d->u, where);
        upload(d, NSB_V, d->v, where);
        upload(d, NSB_OMEGA, d->omega, where);
        NSB_OK(nsb_diagnostics_phys(c, out), where);
    }
}

/* spectral_analysis.c:8-19 */
double ns_spectral_compute_energy(NSSpectralData *d) {
    double s[4];
    sums(d, s, "ns_spectral_compute_energy");
    return 0.5 * (s[0] + s[1]) / (d->nx * d->ny);
}

/* spectral_analysis.c:22-33 */
double ns_spectral_compute_enstrophy(NSSpectralData *d) {
    double s[4];
    sums(d, s, "ns_spectral_compute_enstrophy");
    return 0.5 * s[2] / (d->nx * d->ny);
}

static void lazy_fetch_spectra(NSSpectralData *d, int with_velocity, const char *where) {
    if (!lazy(d)) return;
    download(d, NSB_OMEGA_HAT, d->omega_hat, where);
    if (with_velocity) {
        NSB_OK(nsb_velocity_from_vorticity(ctx_ro(d, where)), where);
        download(d, NSB_U_HAT, d->u_hat, where);
        download(d, NSB_V_HAT, d->v_hat, where);
    }
}

/* spectral_analysis.c:36-66 */
/* reductions over omega_hat for the two driver-loop diagnostics: on the device in LAZY mode (no
 * field download), over the host array otherwise */
/* a device group that has stepped since the last gather: reductions over omega_hat run as partial sums on every
 * device (own columns) and are combined here; nothing is gathered into the first device */
static int group_partial(const NSSpectralData *d) {
    const Backend *b = backend_of(d);
    return b && b->ctx && b->policy == NS_SYNC_LAZY && b->ngpus > 1 && !b->gathered && !b->dirty;
}

static void omega_hat_stats(NSSpectralData *d, double out[6], const char *where) {
    if (group_partial(d)) {
        Backend *b = backend_of(d);
        for (int k = 0; k < 6; ++k) out[k] = 0.0;
        for (int r = 0; r < b->ngpus; ++r) {
            double part[6];
            NSB_OK(nsb_spectral_stats(b->group[r], part), where);
            out[0] += part[0]; out[1] += part[1]; out[3] += part[3]; out[4] += part[4];
            if (part[2] > out[2]) out[2] = part[2];
            out[5] = part[5];
        }
        return;
    }
    if (lazy(d)) {
        NSB_OK(nsb_spectral_stats(ctx_ro(d, where), out), where);
        return;
    }
    const size_t m = n_spec(d);
    const double k2_last = d->k2[m - 1];
    const double max_k = sqrt(k2_last);
    double total = 0.0, high = 0.0, peak = 0.0, tail = 0.0;
    long cnt = 0;
    for (size_t k = 0; k < m; ++k) {
        const double a = cabs(d->omega_hat[k]);
        const double mag2 = a * a;
        total += mag2;
        const double kn = (max_k > 0.0) ? sqrt(d->k2[k]) / max_k : 0.0;
        if (kn > 0.8) high += mag2;
        if (a > peak) peak = a;
        if (d->k2[k] > 0.25 * k2_last) {
            tail += mag2;
            ++cnt;
        }
    }
    out[0] = total; out[1] = high; out[2] = peak; out[3] = tail; out[4] = (double)cnt; out[5] = k2_last;
}

int ns_spectral_check_resolution(NSSpectralData *d) {
    static const char *W = "ns_spectral_check_resolution";
    if (!group_partial(d)) ctx_ro(d, W);
    double st[6];
    omega_hat_stats(d, st, W);
    const double frac = (st[0] > 1e-16) ? st[1] / st[0] : 0.0;
    if (frac > 1e-6) {
        printf("Warning: %.2e%% of energy in high wavenumbers\n", frac * 100);
        return 0;
    }
    return 1;
}

/* spectral_analysis.c:69-100 */
void ns_spectral_analyze_convergence(NSSpectralData *d, int step) {
    static const char *W = "ns_spectral_analyze_convergence";
    if (step % 10 != 0) return;
    if (!group_partial(d)) ctx_ro(d, W);
    double s[4];
    sums(d, s, W);
    const double energy = 0.5 * (s[0] + s[1]) / (d->nx * d->ny);
    const double enstrophy = 0.5 * s[2] / (d->nx * d->ny);
    double st[6];
    omega_hat_stats(d, st, W);
    const double err = st[4] > 0.0 ? sqrt(st[3] / st[4]) : 0.0;
    printf("Step %d: t=%.4f, E=%.6e, Ω=%.6e, max|ω̂|=%.2e, err=%.2e, dt=%.2e\n", step, d->t, energy, enstrophy,
           st[2], err, d->dt);
}

/* spectral_analysis.c:103-142. LAZY: shell sums on the device (nsb_energy_spectrum), nothing is
 * downloaded; EAGER: over the host u_hat, v_hat like the reference. */
int ns_spectral_energy_spectrum(NSSpectralData *d, int nbins, double *ek_sum, int *count) {
    static const char *W = "ns_spectral_energy_spectrum";
    if (!d || nbins <= 0 || !ek_sum || !count) return -1;
    if (group_partial(d)) {
        Backend *b = backend_of(d);
        double *ek = (double *)malloc((size_t)nbins * sizeof(double));
        int *cn = (int *)malloc((size_t)nbins * sizeof(int));
        if (!ek || !cn) { free(ek); free(cn); return -1; }
        memset(ek_sum, 0, (size_t)nbins * sizeof(double));
        memset(count, 0, (size_t)nbins * sizeof(int));
        for (int r = 0; r < b->ngpus; ++r) {
            NSB_OK(nsb_energy_spectrum(b->group[r], nbins, ek, cn), W);
            for (int k = 0; k < nbins; ++k) { ek_sum[k] += ek[k]; count[k] += cn[k]; }
        }
        free(ek);
        free(cn);
        return 0;
    }
    if (lazy(d)) {
        NSB_OK(nsb_energy_spectrum(ctx_ro(d, W), nbins, ek_sum, count), W);
        return 0;
    }
    memset(ek_sum, 0, (size_t)nbins * sizeof(double));
    memset(count, 0, (size_t)nbins * sizeof(int));
    for (size_t k = 0; k < n_spec(d); ++k) {
        const int bin = (int)(sqrt(d->k2[k]) + 0.5);
        if (bin < nbins) {
            const double a = cabs(d->u_hat[k]), b = cabs(d->v_hat[k]);
            ek_sum[bin] += 0.5 * (a * a + b * b);
            count[bin]++;
        }
    }
    return 0;
}

void ns_spectral_analyze_spectrum(NSSpectralData *d) {
    static const char *W = "ns_spectral_analyze_spectrum";
    if (!group_partial(d)) ctx_ro(d, W);
    const int nkx = d->nkx, nky = d->nky;
    const int max_k = (int)fmin(nkx, nky / 2);
    double *ek = (double *)calloc((size_t)max_k, sizeof(double));
    int *cnt = (int *)calloc((size_t)max_k, sizeof(int));
    if (!ek || !cnt || ns_spectral_energy_spectrum(d, max_k, ek, cnt) != 0) {
        free(ek);
        free(cnt);
        return;
    }
    printf("\nEnergy Spectrum (t=%.4f):\n", d->t);
    printf("k\tE(k)\t\tSlope\n");
    for (int k = 1; k < max_k && cnt[k] > 0; ++k) {
        ek[k] /= cnt[k];
        double slope = 0.0;
        if (k > 1 && ek[k - 1] > 1e-16 && ek[k] > 1e-16) slope = log(ek[k] / ek[k - 1]) / log((double)k / (k - 1));
        if (k <= 10 || k % 5 == 0) printf("%d\t%.2e\t%.2f\n", k, ek[k], slope);
    }
    free(ek);
    free(cnt);
}

/* ---------------------------------------------------------- hyperviscosity */

static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* hyperviscosity.c:15-30 */
void ns_spectral_hyperviscosity_default(NSSpectralHyperviscosityConfig *c) {
    if (!c) return;
    memset(c, 0, sizeof *c);
    c->order = 8;
    c->cutoff_ratio = 0.65;
    c->sfd_timescale = 1.0;
}

/* hyperviscosity.c:60-102 (host table kept for callers that read it; the kernels
 * evaluate the same formula per mode) */
void ns_spectral_hyperviscosity_update_weights(NSSpectralData *d) {
    if (!d || !d->hyperviscosity_weights || !d->k2) return;
    const size_t m = n_spec(d);
    NSSpectralHyperviscosityConfig *h = &d->hyperviscosity;
    double kmax = 0.0;
    for (size_t k = 0; k < m; ++k) kmax = fmax(kmax, sqrt(d->k2[k]));
    h->kmax = kmax;
    h->cutoff_wavenumber = clampd(h->cutoff_ratio, 0.0, 1.0) * kmax;
    const double bw = fmax(kmax - h->cutoff_wavenumber, 1e-12);
    h->inv_bandwidth = (bw > 0.0) ? 1.0 / bw : 0.0;
    if (!h->enabled || h->coefficient <= 0.0 || kmax <= 0.0) {
        memset(d->hyperviscosity_weights, 0, m * sizeof(double));
    } else {
        const int order = h->order > 0 ? h->order : 1;
        for (size_t k = 0; k < m; ++k) {
            const double kmag = sqrt(d->k2[k]);
            d->hyperviscosity_weights[k] =
                (kmag > h->cutoff_wavenumber) ? h->coefficient * pow((kmag - h->cutoff_wavenumber) * h->inv_bandwidth, order) : 0.0;
        }
    }
    push_hyperviscosity(d);
}

/* hyperviscosity.c:32-58 */
void ns_spectral_hyperviscosity_set(NSSpectralData *d, const NSSpectralHyperviscosityConfig *cfg) {
    if (!d || !cfg) return;
    const size_t m = n_spec(d);
    d->hyperviscosity = *cfg;
    NSSpectralHyperviscosityConfig *h = &d->hyperviscosity;
    if (!h->enabled) memset(d->hyperviscosity_weights, 0, m * sizeof(double));
    if (!h->sfd_enabled || h->sfd_strength <= 0.0) {
        h->sfd_enabled = 0;
        h->sfd_strength = 0.0;
        memset(d->sfd_hat, 0, m * sizeof(fftw_complex));
    }
    ns_spectral_hyperviscosity_update_weights(d);
    if (h->sfd_enabled) {
        Backend *b = backend_of(d);
        if (b && b->ctx && lazy(d)) {
            /* through ctx_of: on a device group the columns of all devices are gathered first and the
             * seeded filter state is published again before the next group step */
            NSB_OK(nsb_copy(ctx_of(d, "ns_spectral_hyperviscosity_set"), NSB_SFD_HAT, NSB_OMEGA_HAT),
                   "ns_spectral_hyperviscosity_set");
        } else {
            memcpy(d->sfd_hat, d->omega_hat, m * sizeof(fftw_complex));
            if (b && b->ctx) upload(d, NSB_SFD_HAT, d->sfd_hat, "ns_spectral_hyperviscosity_set");
        }
    }
}

/* hyperviscosity.c:104-150 */
void ns_spectral_hyperviscosity_apply(NSSpectralData *d, fftw_complex *rhs_hat, const fftw_complex *state_hat) {
    static const char *W = "ns_spectral_hyperviscosity_apply";
    if (!d || !rhs_hat || !state_hat) return;
    const NSSpectralHyperviscosityConfig *h = &d->hyperviscosity;
    const int visc = (h->enabled && h->coefficient > 0.0), sfd = (h->sfd_enabled && h->sfd_strength > 0.0);
    if (!visc && !sfd) return;
    nsb_ctx *c = ctx_of(d, W);
    push_hyperviscosity(d);
    upload(d, NSB_SCRATCH_HAT, rhs_hat, W);
    upload(d, NSB_SCRATCH2_HAT, state_hat, W);
    if (sfd) upload(d, NSB_SFD_HAT, d->sfd_hat, W);
    NSB_OK(nsb_hyperviscosity_apply(c, NSB_SCRATCH_HAT, NSB_SCRATCH2_HAT), W);
    download(d, NSB_SCRATCH_HAT, rhs_hat, W);
}

/* hyperviscosity.c:152-178 */
void ns_spectral_hyperviscosity_post_step(NSSpectralData *d, const fftw_complex *state_hat, double dt) {
    static const char *W = "ns_spectral_hyperviscosity_post_step";
    if (!d || !state_hat || dt <= 0.0) return;
    const NSSpectralHyperviscosityConfig *h = &d->hyperviscosity;
    if (!h->sfd_enabled || h->sfd_strength <= 0.0 || h->sfd_timescale <= 0.0) return;
    nsb_ctx *c = ctx_of(d, W);
    push_hyperviscosity(d);
    upload(d, NSB_SCRATCH2_HAT, state_hat, W);
    upload(d, NSB_SFD_HAT, d->sfd_hat, W);
    NSB_OK(nsb_hyperviscosity_post_step(c, NSB_SCRATCH2_HAT, dt), W);
    download(d, NSB_SFD_HAT, d->sfd_hat, W);
}

/* ------------------------------------------------------------------ output */

/* spectral_output.c:4-27 with every `stride`-th point in x and y (stride 1 = the reference's file,
 * byte for byte the same format: scripts/plot_solution.py and energy_spectrum.py read it). A full
 * ASCII dump is 76 bytes per point, 20 GB at 16384^2, hence the stride. */
int ns_spectral_output_solution_strided(NSSpectralData *d, const char *filename, int stride) {
    static const char *W = "ns_spectral_output_solution";
    nsb_ctx *c = ctx_ro(d, W);
    if (!filename || stride < 1) return -1;
    FILE *fp = fopen(filename, "w");
    if (!fp) return -1;
    if (lazy(d)) {
        NSB_OK(nsb_compute_derivatives(c), W);
        download(d, NSB_U, d->u, W);
        download(d, NSB_V, d->v, W);
        download(d, NSB_OMEGA, d->omega, W);
    } else {
        upload(d, NSB_PSI_HAT, d->psi_hat, W);
    }
    const int src = NSB_PSI_HAT, dst = NSB_PSI;
    NSB_OK(nsb_spectral_to_physical(c, 1, &src, &dst), W);
    download(d, NSB_PSI, d->psi, W);
    fprintf(fp, "# Spectral Navier-Stokes Solution\n");
    fprintf(fp, "# x y u v omega psi\n");
    for (int j = 0; j < d->ny; j += stride) {
        for (int i = 0; i < d->nx; i += stride) {
            const size_t idx = (size_t)j * d->nx + i;
            fprintf(fp, "%.8f %.8f %.8e %.8e %.8e %.8e\n", d->x[i], d->y[j], d->u[idx], d->v[idx], d->omega[idx],
                    d->psi[idx]);
        }
        fprintf(fp, "\n");
    }
    return fclose(fp) == 0 ? 0 : -1;
}

/* spectral_output.c:4-27 */
void ns_spectral_output_solution(NSSpectralData *d, const char *filename) {
    (void)ns_spectral_output_solution_strided(d, filename, 1);
}

/* -------------------------------------------------------------- restartable state */

/*
 * Raw binary restart file. The reference's HDF5 checkpoint
 * (src/common/checkpoint.c:266-301) stores u, v, omega, psi and three spectra but neither the SFD
 * state nor the forcing stream, so a forced run cannot resume bit for bit from it. Here exactly the
 * state the time loop owns is kept: omega_hat, sfd_hat (if the filter is on), t, dt, and of the
 * forcing generator rng_state, have_spare, spare_gaussian and the coloured-noise state. A resumed
 * run reproduces the uninterrupted one bit for bit (tests/test_host_api.py).
 * Layout: NSRestartHeader | omega_hat[M] | sfd_hat[M] if has_sfd | colored_state[M] if has_colored.
 */
typedef struct {
    char magic[8]; /* "NSBRST1\0" */
    int nx, ny;
    double Lx, Ly, Re, t, dt, cfl_factor;
    NSSpectralHyperviscosityConfig hv;
    int has_sfd, has_forcing, has_colored, pad_;
    NSForcingConfig fcfg;
    unsigned long long rng_state;
    double spare_gaussian;
    unsigned int have_spare;
    int active_modes;
} NSRestartHeader;

static int write_all(FILE *fp, const void *p, size_t bytes) {
    const char *q = (const char *)p;
    while (bytes > 0) { /* chunks: a single fwrite of > 2 GB is not portable */
        const size_t n = bytes > ((size_t)1 << 30) ? ((size_t)1 << 30) : bytes;
        if (fwrite(q, 1, n, fp) != n) return -1;
        q += n;
        bytes -= n;
    }
    return 0;
}

static int read_all(FILE *fp, void *p, size_t bytes) {
    char *q = (char *)p;
    while (bytes > 0) {
        const size_t n = bytes > ((size_t)1 << 30) ? ((size_t)1 << 30) : bytes;
        if (fread(q, 1, n, fp) != n) return -1;
        q += n;
        bytes -= n;
    }
    return 0;
}

/* <filename>.rank<r>: NSSlabHeader | omega_hat[ny][c1-c0] | sfd_hat[ny][c1-c0] if has_sfd */
typedef struct {
    char magic[8]; /* "NSBSLB1\0" */
    int rank, nranks, col0, col1, ny, has_sfd;
} NSSlabHeader;

static int save_slab_file(NSSpectralData *d, nsb_ctx *c, int rank, int nranks, int sfd, const char *filename) {
    char path[1024];
    snprintf(path, sizeof path, "%s.rank%d", filename, rank);
    const size_t cap = ((size_t)d->nkx / (size_t)nranks + 16) * (size_t)d->nky * sizeof(fftw_complex);
    void *buf = malloc(cap);
    if (!buf) return -1;
    NSSlabHeader sh;
    memset(&sh, 0, sizeof sh);
    memcpy(sh.magic, "NSBSLB1", 8);
    sh.rank = rank; sh.nranks = nranks; sh.ny = d->nky; sh.has_sfd = sfd;
    int rc = nsb_download_columns(c, NSB_OMEGA_HAT, buf, &sh.col0, &sh.col1);
    FILE *fp = rc == 0 ? fopen(path, "wb") : NULL;
    if (!fp) { free(buf); return -1; }
    const size_t bytes = (size_t)(sh.col1 - sh.col0) * (size_t)d->nky * sizeof(fftw_complex);
    rc = write_all(fp, &sh, sizeof sh);
    if (rc == 0) rc = write_all(fp, buf, bytes);
    if (rc == 0 && sfd) {
        rc = nsb_download_columns(c, NSB_SFD_HAT, buf, NULL, NULL);
        if (rc == 0) rc = write_all(fp, buf, bytes);
    }
    if (fclose(fp) != 0) rc = -1;
    free(buf);
    return rc;
}

/* the slab files of a per-device restart back into the host arrays (any device count may resume from them) */
static int load_slab_files(NSSpectralData *d, int nranks, int has_sfd, const char *filename) {
    memset(d->omega_hat, 0, n_spec(d) * sizeof(fftw_complex));
    if (has_sfd) memset(d->sfd_hat, 0, n_spec(d) * sizeof(fftw_complex));
    for (int r = 0; r < nranks; ++r) {
        char path[1024];
        snprintf(path, sizeof path, "%s.rank%d", filename, r);
        FILE *fp = fopen(path, "rb");
        if (!fp) return -1;
        NSSlabHeader sh;
        int rc = read_all(fp, &sh, sizeof sh);
        if (rc != 0 || memcmp(sh.magic, "NSBSLB1", 8) != 0 || sh.rank != r || sh.nranks != nranks || sh.ny != d->nky ||
            sh.col0 < 0 || sh.col1 > d->nkx || sh.col1 < sh.col0 || (has_sfd && !sh.has_sfd)) {
            fclose(fp);
            return -1;
        }
        const int w = sh.col1 - sh.col0;
        fftw_complex *row = (fftw_complex *)malloc((size_t)(w > 0 ? w : 1) * sizeof(fftw_complex));
        if (!row) { fclose(fp); return -1; }
        for (int pass = 0; pass < (has_sfd ? 2 : 1) && rc == 0; ++pass) {
            fftw_complex *dst = pass ? d->sfd_hat : d->omega_hat;
            for (int j = 0; j < d->nky && rc == 0; ++j) {
                rc = read_all(fp, row, (size_t)w * sizeof(fftw_complex));
                if (rc == 0) memcpy(dst + (size_t)j * d->nkx + sh.col0, row, (size_t)w * sizeof(fftw_complex));
            }
        }
        free(row);
        fclose(fp);
        if (rc != 0) return -1;
    }
    return 0;
}

int ns_spectral_save_state(NSSpectralData *d, const NSForcingState *fs, const char *filename) {
    static const char *W = "ns_spectral_save_state";
    if (!d || !filename) { set_error("%s: NULL argument", W); return -1; }
    Backend *bk = backend_of(d);
    if (!bk || !bk->ctx) { set_error("%s: no device context", W); return -1; }
    const NSSpectralHyperviscosityConfig *h = &d->hyperviscosity;
    const int sfd = (h->sfd_enabled && h->sfd_strength > 0.0);
    /* a slab-decomposed group that has stepped since the last gather: every device writes its own column
     * block to <filename>.rank<r> (per-GPU slab files), nothing funnels through device 0 */
    const int per_device = lazy(d) && bk->ngpus > 1 && !bk->gathered && !bk->dirty;
    if (!per_device) ctx_ro(d, W);
    if (lazy(d)) { /* the device copy is the truth */
        if (!per_device) {
            download(d, NSB_OMEGA_HAT, d->omega_hat, W);
            if (sfd) download(d, NSB_SFD_HAT, d->sfd_hat, W);
        }
        int steps;
        NSB_OK(nsb_get_state(bk->ctx, &d->dt, &d->t, &steps), W);
    }
    NSRestartHeader hd;
    memset(&hd, 0, sizeof hd);
    memcpy(hd.magic, "NSBRST1", 8);
    hd.nx = d->nx; hd.ny = d->ny;
    hd.Lx = d->Lx; hd.Ly = d->Ly; hd.Re = d->Re; hd.t = d->t; hd.dt = d->dt; hd.cfl_factor = d->cfl_factor;
    /* field by field: struct padding must not leak stack bytes into the file (identical states give
     * identical files) */
    hd.hv.enabled = h->enabled; hd.hv.coefficient = h->coefficient; hd.hv.order = h->order;
    hd.hv.cutoff_ratio = h->cutoff_ratio; hd.hv.sfd_strength = h->sfd_strength;
    hd.hv.sfd_timescale = h->sfd_timescale; hd.hv.sfd_enabled = h->sfd_enabled; hd.hv.kmax = h->kmax;
    hd.hv.cutoff_wavenumber = h->cutoff_wavenumber; hd.hv.inv_bandwidth = h->inv_bandwidth;
    hd.has_sfd = sfd;
    if (fs) {
        if (fs->nkx != d->nkx || fs->nky != d->nky) { set_error("%s: forcing state does not match the grid", W); return -1; }
        hd.has_forcing = 1;
        hd.has_colored = (fs->config.type == NS_FORCING_COLORED_NOISE && fs->colored_state != NULL);
        hd.fcfg.kmin = fs->config.kmin; hd.fcfg.kmax = fs->config.kmax;
        hd.fcfg.energy_target = fs->config.energy_target;
        hd.fcfg.correlation_time = fs->config.correlation_time;
        hd.fcfg.seed = fs->config.seed; hd.fcfg.type = fs->config.type;
        hd.rng_state = fs->rng_state;
        hd.spare_gaussian = fs->spare_gaussian;
        hd.have_spare = fs->have_spare;
        hd.active_modes = fs->active_modes;
    }
    hd.pad_ = per_device ? bk->ngpus : 0; /* number of per-device slab files that hold the fields (0: in this file) */
    FILE *fp = fopen(filename, "wb");
    if (!fp) { set_error("%s: cannot open %s", W, filename); return -1; }
    const size_t mb = n_spec(d) * sizeof(fftw_complex);
    int rc = write_all(fp, &hd, sizeof hd);
    if (per_device) {
        for (int r = 0; r < bk->ngpus && rc == 0; ++r) rc = save_slab_file(d, bk->group[r], r, bk->ngpus, sfd, filename);
    } else {
        if (rc == 0) rc = write_all(fp, d->omega_hat, mb);
        if (rc == 0 && hd.has_sfd) rc = write_all(fp, d->sfd_hat, mb);
    }
    if (rc == 0 && hd.has_colored) rc = write_all(fp, fs->colored_state, mb);
    if (fclose(fp) != 0) rc = -1;
    if (rc != 0) set_error("%s: short write to %s", W, filename);
    return rc;
}

int ns_spectral_load_state(NSSpectralData *d, NSForcingState *fs, const char *filename) {
    static const char *W = "ns_spectral_load_state";
    if (!d || !filename) { set_error("%s: NULL argument", W); return -1; }
    nsb_ctx *c = ctx_of(d, W);
    FILE *fp = fopen(filename, "rb");
    if (!fp) { set_error("%s: cannot open %s", W, filename); return -1; }
    NSRestartHeader hd;
    if (read_all(fp, &hd, sizeof hd) != 0 || memcmp(hd.magic, "NSBRST1", 8) != 0) {
        fclose(fp);
        set_error("%s: %s is not a restart file", W, filename);
        return -1;
    }
    if (hd.nx != d->nx || hd.ny != d->ny || hd.Lx != d->Lx || hd.Ly != d->Ly) {
        fclose(fp);
        set_error("%s: file holds a %d x %d grid on %g x %g, the solver has %d x %d on %g x %g", W, hd.nx, hd.ny, hd.Lx,
                  hd.Ly, d->nx, d->ny, d->Lx, d->Ly);
        return -1;
    }
    if (hd.Re != d->Re) {
        fclose(fp);
        set_error("%s: file was written at Re = %.17g, the solver runs at Re = %.17g", W, hd.Re, d->Re);
        return -1;
    }
    if (fs && hd.has_forcing &&
        (fs->nkx != d->nkx || fs->nky != d->nky || fs->active_modes != hd.active_modes ||
         fs->config.type != hd.fcfg.type || fs->config.kmin != hd.fcfg.kmin || fs->config.kmax != hd.fcfg.kmax ||
         (hd.has_colored && !fs->colored_state))) {
        fclose(fp);
        set_error("%s: forcing state was initialised with a different band or type than the file's", W);
        return -1;
    }
    const size_t mb = n_spec(d) * sizeof(fftw_complex);
    int rc = 0;
    if (hd.pad_ > 1) { /* fields live in per-device slab files */
        rc = load_slab_files(d, hd.pad_, hd.has_sfd, filename);
        if (rc != 0) { fclose(fp); set_error("%s: cannot read the %d slab files %s.rank*", W, hd.pad_, filename); return -1; }
    } else {
        rc = read_all(fp, d->omega_hat, mb);
        if (rc == 0 && hd.has_sfd) rc = read_all(fp, d->sfd_hat, mb);
    }
    if (rc == 0 && hd.has_colored) {
        if (fs && hd.has_forcing) rc = read_all(fp, fs->colored_state, mb);
    }
    fclose(fp);
    if (rc != 0) { set_error("%s: %s is truncated", W, filename); return -1; }
    d->t = hd.t; d->dt = hd.dt; d->cfl_factor = hd.cfl_factor;
    /* hyperviscosity_set would reseed sfd_hat from omega_hat (hyperviscosity.c:53-56): restore the
     * configuration and weights without touching the filter state */
    d->hyperviscosity = hd.hv;
    ns_spectral_hyperviscosity_update_weights(d);
    upload(d, NSB_OMEGA_HAT, d->omega_hat, W);
    if (hd.has_sfd) upload(d, NSB_SFD_HAT, d->sfd_hat, W);
    NSB_OK(nsb_set_time(c, d->t), W);
    push_physics(d, W); /* the file's cfl_factor is live on the device(s) from here on */
    if (fs && hd.has_forcing) {
        fs->config = hd.fcfg;
        fs->rng_state = hd.rng_state;
        fs->spare_gaussian = hd.spare_gaussian;
        fs->have_spare = hd.have_spare;
    }
    if (!lazy(d)) refresh_derived_to_host(d, W);
    return 0;
}

/* ---------------------------------------------------------------- additive */

int ns_spectral_backend_select_device(NSSpectralData *d, int device) {
    Backend *b = backend_of(d);
    if (!b || b->ctx) return -1;
    b->device = device;
    return 0;
}

/* slab-decompose this solver over n_gpus devices driven by this process (devices
 * device .. device+n_gpus-1; NS_SPECTRAL_GROUP_SAME_DEVICE=1 keeps them all on `device`, for
 * machines with fewer GPUs). Call after ns_spectral_initialize_fft_plans. ns_spectral_run_steps
 * then advances the one grid on all devices; every other entry point gathers the state into the
 * first device and runs there. */
int ns_spectral_backend_init(NSSpectralData *d, int n_gpus) {
    static const char *W = "ns_spectral_backend_init";
    Backend *b = backend_of(d);
    if (!b || !b->ctx) { set_error("%s: call ns_spectral_initialize_fft_plans first", W); return -1; }
    if (b->ngpus > 1) { set_error("%s: already initialised", W); return -1; }
    if (n_gpus <= 1) return 0;
    if (n_gpus > 8) { set_error("%s: at most 8 devices", W); return -1; }
    const char *same = getenv("NS_SPECTRAL_GROUP_SAME_DEVICE");
    const int same_dev = same && atoi(same);
    b->group[0] = b->ctx;
    for (int r = 1; r < n_gpus; ++r) {
        nsb_params p;
        memset(&p, 0, sizeof p);
        p.nx = d->nx; p.ny = d->ny; p.Lx = d->Lx; p.Ly = d->Ly; p.Re = d->Re;
        p.cfl_factor = d->cfl_factor;
        p.device = same_dev ? b->device : b->device + r;
        if (nsb_ctx_create(&p, &b->group[r]) != 0) {
            set_error("%s: %s", W, nsb_last_error());
            for (int q = 1; q < r; ++q) nsb_ctx_destroy(b->group[q]);
            return -1;
        }
    }
    for (int r = 0; r < n_gpus; ++r) {
        if (nsb_mgpu_setup(b->group[r], r, n_gpus) != 0) {
            set_error("%s: %s", W, nsb_last_error());
            for (int q = 1; q < n_gpus; ++q) nsb_ctx_destroy(b->group[q]);
            nsb_mgpu_setup(b->ctx, 0, 1);
            return -1;
        }
    }
    if (nsb_mgpu_link_local(b->group, n_gpus) != 0) {
        set_error("%s: %s", W, nsb_last_error());
        for (int q = 1; q < n_gpus; ++q) nsb_ctx_destroy(b->group[q]);
        nsb_mgpu_setup(b->ctx, 0, 1);
        return -1;
    }
    b->ngpus = n_gpus;
    b->gathered = 1; /* the first device holds everything so far */
    b->dirty = 1;
    push_hyperviscosity(d);
    return 0;
}

void ns_spectral_set_sync_policy(NSSpectralData *d, NSSyncPolicy policy) {
    Backend *b = backend_of(d);
    if (!b) return;
    if (b->ctx && b->policy == NS_SYNC_EAGER && policy == NS_SYNC_LAZY) {
        /* the device copy becomes the truth: seed it from the host arrays */
        upload(d, NSB_OMEGA_HAT, d->omega_hat, "ns_spectral_set_sync_policy");
        push_sfd_if_needed(d, "ns_spectral_set_sync_policy");
        NSB_OK(nsb_set_time(b->ctx, d->t), "ns_spectral_set_sync_policy");
    } else if (b->ctx && b->policy == NS_SYNC_LAZY && policy == NS_SYNC_EAGER) {
        ns_spectral_sync_to_host(d, NS_FIELD_ALL);
    }
    b->policy = policy;
}

NSSyncPolicy ns_spectral_get_sync_policy(const NSSpectralData *d) {
    const Backend *b = backend_of(d);
    return b ? b->policy : NS_SYNC_EAGER;
}

int ns_spectral_sync_to_host(NSSpectralData *d, unsigned mask) {
    static const char *W = "ns_spectral_sync_to_host";
    nsb_ctx *c = ctx_ro(d, W);
    if (mask & NS_FIELD_OMEGA_HAT) download(d, NSB_OMEGA_HAT, d->omega_hat, W);
    if (mask & NS_FIELD_SFD_HAT) {
        const NSSpectralHyperviscosityConfig *h = &d->hyperviscosity;
        if (h->sfd_enabled && h->sfd_strength > 0.0) download(d, NSB_SFD_HAT, d->sfd_hat, W);
    }
    if (mask & NS_FIELD_DERIVED) {
        NSB_OK(nsb_compute_derivatives(c), W);
        if (mask & NS_FIELD_PSI_HAT) download(d, NSB_PSI_HAT, d->psi_hat, W);
        if (mask & NS_FIELD_U_HAT) download(d, NSB_U_HAT, d->u_hat, W);
        if (mask & NS_FIELD_V_HAT) download(d, NSB_V_HAT, d->v_hat, W);
        if (mask & NS_FIELD_U) download(d, NSB_U, d->u, W);
        if (mask & NS_FIELD_V) download(d, NSB_V, d->v, W);
        if (mask & NS_FIELD_OMEGA) download(d, NSB_OMEGA, d->omega, W);
    }
    if (mask & NS_FIELD_PSI) {
        if (!(mask & NS_FIELD_DERIVED)) NSB_OK(nsb_velocity_from_vorticity(c), W);
        const int src = NSB_PSI_HAT, dst = NSB_PSI;
        NSB_OK(nsb_spectral_to_physical(c, 1, &src, &dst), W);
        download(d, NSB_PSI, d->psi, W);
    }
    double dt;
    int steps;
    NSB_OK(nsb_get_state(c, &dt, &d->t, &steps), W);
    return 0;
}

int ns_spectral_sync_to_device(NSSpectralData *d, unsigned mask) {
    static const char *W = "ns_spectral_sync_to_device";
    nsb_ctx *c = ctx_of(d, W);
    if (mask & NS_FIELD_OMEGA_HAT) upload(d, NSB_OMEGA_HAT, d->omega_hat, W);
    if (mask & NS_FIELD_SFD_HAT) upload(d, NSB_SFD_HAT, d->sfd_hat, W);
    if (mask & NS_FIELD_PSI_HAT) upload(d, NSB_PSI_HAT, d->psi_hat, W);
    NSB_OK(nsb_set_time(c, d->t), W);
    return 0;
}

void *ns_spectral_backend_context(NSSpectralData *d) {
    Backend *b = backend_of(d);
    return b ? (void *)b->ctx : NULL;
}
