/*
 * include/nsb.h -- the thin kernel C-ABI ("nsb" = Navier-Stokes B200).
 *
 * Plain C, POD arguments, opaque device context, int status (0 = ok, <0 = error;
 * text via nsb_last_error()). No torch / C++ types cross this boundary. This is
 * the layer the reference-facing entry points in include/ns_spectral_solver.h are
 * implemented on (navier-stokes-solvers_b200/csrc/host/spectral_api.c); each call
 * names the reference function whose work it performs on the device.
 *
 * All arrays keep the reference's layouts
 * (/root/reference/src/spectral/ns_spectral_solver.h:22-68):
 *   physical  f[j*nx + i]          doubles,          nx*ny
 *   spectral  f_hat[j*nkx + i]     (re,im) doubles,  nkx*ny with nkx = nx/2+1
 * nx, ny must be powers of two >= 8.
 */
#ifndef NSB_H
#define NSB_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nsb_ctx nsb_ctx;

typedef struct {
    int nx, ny;
    double Lx, Ly, Re;
    double cfl_factor; /* reference default 0.3 (spectral_memory.c:19) */
    int device;        /* CUDA device ordinal */
    /* tuning / test overrides, 0 = automatic */
    int y1;            /* ny = y1 * Y2 : length of the in-SM column FFT */
    int x2;            /* nx = X1 * x2 : length of the in-SM row FFT    */
    int col_tile;      /* columns per column-FFT CTA */
    int beta_batch;    /* k2 values per beta CTA */
} nsb_params;

typedef struct {
    int enabled;              /* cfg.enabled && coefficient > 0 */
    double coefficient;
    int order;                /* already clamped to >= 1 */
    double cutoff_wavenumber; /* derived by ns_spectral_hyperviscosity_update_weights */
    double inv_bandwidth;
    int sfd_enabled;          /* cfg.sfd_enabled && sfd_strength > 0 */
    double sfd_strength;
    double sfd_timescale;
} nsb_hyperviscosity;

/* device-resident fields (allocated on first use) */
enum {
    NSB_OMEGA_HAT = 0, /* the RK state */
    NSB_PSI_HAT = 1,
    NSB_U_HAT = 2,
    NSB_V_HAT = 3,
    NSB_NONLINEAR_HAT = 4,
    NSB_SFD_HAT = 5,
    NSB_RK_ACC = 6,      /* k1 + 2 k2 + 2 k3 accumulator */
    NSB_OMEGA_STAGE = 7, /* stage state omega_0 + c dt k */
    NSB_SCRATCH_HAT = 8,
    NSB_SCRATCH2_HAT = 9,
    NSB_NUM_SPECTRAL = 10,
    NSB_U = 16,
    NSB_V = 17,
    NSB_P = 18,
    NSB_OMEGA = 19,
    NSB_PSI = 20,
    NSB_SCRATCH_PHYS = 21,
    NSB_NUM_FIELDS = 22
};

const char *nsb_last_error(void);
int nsb_device_count(void);

/* ns_spectral_allocate + initialize_grids + initialize_fft_plans (spectral_memory.c:6,
 * spectral_initialization.c:10,57): wavenumber + twiddle tables, work buffers, stream. */
int nsb_ctx_create(const nsb_params *p, nsb_ctx **out);
void nsb_ctx_destroy(nsb_ctx *c);
int nsb_sync(nsb_ctx *c);

/* host <-> device copies of one whole field (reference layout). Synchronous. */
int nsb_upload(nsb_ctx *c, int field, const void *host);
int nsb_download(nsb_ctx *c, int field, void *host);
int nsb_zero(nsb_ctx *c, int field);
/* the columns [col0, col1) this device owns (all of the retained ones on a single device) of a spectral field as a
 * dense ny x (col1 - col0) host block: what a per-device restart file of a slab-decomposed run holds */
int nsb_download_columns(nsb_ctx *c, int field, void *host, int *col0, int *col1);
int nsb_copy(nsb_ctx *c, int dst_field, int src_field);

/* ns_spectral_hyperviscosity_set (hyperviscosity.c:32-58): parameters only; weights are
 * recomputed from (kx,ky) inside the kernels. */
int nsb_set_hyperviscosity(nsb_ctx *c, const nsb_hyperviscosity *h);

/* Forcing for the NEXT nsb_rk4_step calls (k_s += f_hat, constant over the
 * step). box[(band+jb)*(ib+1)+i] complex, i in [0,ib], band in [-jb,jb]; copied
 * asynchronously through a pinned ring. box == NULL disables forcing. */
int nsb_set_forcing(nsb_ctx *c, int ib, int jb, const double *box);

/* ns_spectral_physical_to_spectral (spectral_transforms.c:9-22): r2c, 1/(nx ny), mask */
int nsb_physical_to_spectral(nsb_ctx *c, int src_phys, int dst_hat);
/* ns_spectral_spectral_to_physical (spectral_transforms.c:25-32): mask, c2r; up to 4 fields */
int nsb_spectral_to_physical(nsb_ctx *c, int nf, const int *src_hat, const int *dst_phys);
/* ns_spectral_apply_dealiasing (spectral_transforms.c:35-43) */
int nsb_apply_dealiasing(nsb_ctx *c, int field_hat);
/* ns_spectral_compute_velocity_from_vorticity (spectral_core.c:8-34) */
int nsb_velocity_from_vorticity(nsb_ctx *c);
/* ns_spectral_compute_derivatives (spectral_core.c:37-43): + physical u, v, omega */
int nsb_compute_derivatives(nsb_ctx *c);
/* ns_spectral_compute_nonlinear_terms (spectral_core.c:46-112): NONLINEAR_HAT from the
 * stored PSI_HAT and OMEGA_HAT */
int nsb_nonlinear_terms(nsb_ctx *c);
/* ns_spectral_hyperviscosity_apply / _post_step (hyperviscosity.c:104-178) */
int nsb_hyperviscosity_apply(nsb_ctx *c, int rhs_hat, int state_hat);
int nsb_hyperviscosity_post_step(nsb_ctx *c, int state_hat, double dt);

/* ns_spectral_compute_cfl_timestep (spectral_time_integration.c:9-31) from the device
 * copies of the physical u, v arrays. */
int nsb_cfl_timestep_uv(nsb_ctx *c, double *dt_out);
/* same, but u, v are formed on the fly from OMEGA_HAT and never stored */
int nsb_cfl_timestep(nsb_ctx *c, double *dt_out);

/* ns_spectral_advance_rk4 (spectral_time_integration.c:34-106) on OMEGA_HAT, device
 * resident, asynchronous. dt > 0: use it (data->dt semantics). dt <= 0: the CFL step is
 * computed on the device from the stage-1 velocity field (the fused max|u| reduction) and
 * never visits the host. */
int nsb_rk4_step(nsb_ctx *c, double dt);

/* The same step in two halves with the device's CFL step size returned in between (8 bytes device -> host), for
 * coloured forcing, whose generator needs dt before it consumes its stream (src/common/forcing.c:138-147):
 * begin runs stage 1's transforms, the fused max |u| and the CFL kernel and returns dt; the caller sets the forcing;
 * finish runs the RK updates and stages 2-4. ctxs: one context, or the contexts of a single-process device group.
 * Fast path only (grids with Y1 == X2 == 1024): otherwise begin returns < 0 and nothing has been launched. */
int nsb_rk4_step_begin(nsb_ctx **ctxs, int n, double *dt_out);
int nsb_rk4_step_finish(nsb_ctx **ctxs, int n);

/* out[0..2] = sum u^2, sum v^2, sum omega^2 over the grid, out[3] = max(u^2+v^2), from
 * OMEGA_HAT (ns_spectral_compute_energy/_enstrophy, spectral_analysis.c:8-33, without
 * materialising the fields). Synchronous. */
int nsb_diagnostics(nsb_ctx *c, double out[4]);
/* same sums from the device copies of the physical arrays U, V, OMEGA */
int nsb_diagnostics_phys(nsb_ctx *c, double out[4]);
/* out[0..2] as nsb_diagnostics, by Parseval over the retained modes (Hermitian projection of the kx = 0 column, as
 * the c2r of spectral_transforms.c:25-32 applies it) instead of three c2r transforms; on a context of a
 * slab-decomposed group over the own columns only: add the results of all devices. Synchronous. */
int nsb_diagnostics_spectral(nsb_ctx *c, double out[3]);

/* The reductions over OMEGA_HAT that ns_spectral_check_resolution (spectral_analysis.c:36-66) and
 * ns_spectral_analyze_convergence (:69-100) make, on the device: out = { sum |w|^2, sum |w|^2 over
 * sqrt(k2)/sqrt(k2_last) > 0.8, max |w|, sum |w|^2 over k2 > 0.25 k2_last, number of such modes,
 * k2_last } with k2_last = k2[nkx*nky - 1] as in the reference. Synchronous. On a context of a slab-decomposed
 * group the sums run over the device's own columns only: add the sums (take the max of out[2]) of all devices. */
int nsb_spectral_stats(nsb_ctx *c, double out[6]);

/* ns_spectral_analyze_spectrum (spectral_analysis.c:103-142) without downloading omega_hat, u_hat,
 * v_hat: ek_sum[b] = sum of 0.5 (|u_hat|^2 + |v_hat|^2) and count[b] = number of modes over the
 * shell b = (int)(sqrt(k2) + 0.5), b < nbins (host arrays of nbins entries). Synchronous. Own columns only on a
 * context of a slab-decomposed group (add the arrays of all devices). */
int nsb_energy_spectrum(nsb_ctx *c, int nbins, double *ek_sum, int *count);

/* Re and cfl_factor are read live from NSSpectralData by the reference (spectral_time_integration.c:
 * 22-24, 52); the host layer pushes the caller's current values before every step / CFL evaluation. */
int nsb_set_physics(nsb_ctx *c, double Re, double cfl_factor);

/* max |f| over a physical field, and f *= s, on the device (the two host loops of the build-defined
 * random-phase initial condition, SURVEY 8d, without moving the field to the host). */
int nsb_phys_absmax(nsb_ctx *c, int field_phys, double *out);
int nsb_phys_scale(nsb_ctx *c, int field_phys, double s);

/* device-side scalars: last dt, time, number of steps taken. Synchronous. */
int nsb_get_state(nsb_ctx *c, double *dt, double *t, int *steps);
int nsb_set_time(nsb_ctx *c, double t);
/* 1 if a device-side CFL evaluation (nsb_rk4_step with dt <= 0, nsb_cfl_timestep) has seen a NaN or Inf in
 * max |u|^2 since OMEGA_HAT was last uploaded -- the field has blown up; *step_out = the step counter at that
 * moment (-1 if finite). 0 if finite, < 0 on error. The reference silently continues with dt = MAX_DT
 * (spectral_time_integration.c:14-20 never sees a NaN); SURVEY section 5 asks for the flag. Synchronous. */
int nsb_check_finite(nsb_ctx *c, int *step_out);
/* the last n (<= 256) step sizes, oldest first */
int nsb_get_dt_history(nsb_ctx *c, int n, double *out);

/* the factorisation in use: out = {Y1, Y2, X1, X2, col_tile, beta_batch, ncol, launches_per_step} */
int nsb_get_plan(nsb_ctx *c, int out[8]);
/* cudaStream_t the context launches on (for CUDA-event timing by the caller) */
void *nsb_stream(nsb_ctx *c);
/* CUDA-event stopwatch on the context's own stream: start records an event, stop records
 * a second one, synchronises on it and returns the elapsed device time in ms. */
int nsb_timer_start(nsb_ctx *c);
int nsb_timer_stop(nsb_ctx *c, float *ms);
/* Per-kernel device timing with CUDA events on the context's stream, for the roofline
 * numbers bench.py reports. Between begin and end every transform kernel launch is
 * bracketed by two events; end returns the summed ms and launch counts per kind:
 * 0 alpha_inv, 1 beta_inv, 2 gamma, 3 beta_fwd, 4 alpha_fwd, 5 other. */
int nsb_profile_begin(nsb_ctx *c);
int nsb_profile_end(nsb_ctx *c, double ms[6], int count[6]);
/*
 * Slab decomposition over the devices of one NVSwitch domain (the reference
 * has no distributed code: /root/reference/src/spectral/spectral_initialization.c:58-61 is
 * OpenMP + FFTW threads). Spectral space is split into blocks of retained kx columns,
 * physical space into blocks of y rows; the two global transposes of every 2-D transform
 * are peer-memory stores fused into the producing kernels (no collective library on the
 * data path). Available on the fast path only (grids 2048..16384 with Y1 == X2 == 1024).
 *   nsb_mgpu_setup       partition for (rank, world); call before nsb_upload
 *   nsb_mgpu_export      one process per device: 3 x 64-byte CUDA IPC handles (T, T', scalars)
 *   nsb_mgpu_import      ... of peer `peer`, as gathered by the caller (torch.distributed, MPI, ...)
 *   nsb_mgpu_link_local  one process driving all devices: enable peer access, swap pointers
 *   nsb_mgpu_group_step  ns_spectral_advance_rk4 on such a group (nsb_rk4_step is used
 *                        per rank in the one-process-per-device mode)
 *   nsb_mgpu_partition   out = {rank, world, col0, col1, row0, rows}
 * nsb_upload / nsb_download keep the reference layout; after steps only the own columns
 * [col0, col1) of OMEGA_HAT are current on each device.
 */
int nsb_mgpu_setup(nsb_ctx *c, int rank, int world);
int nsb_mgpu_partition(nsb_ctx *c, int out[6]);
int nsb_mgpu_export(nsb_ctx *c, unsigned char *handles);
int nsb_mgpu_import(nsb_ctx *c, int peer, const unsigned char *handles);
int nsb_mgpu_link_local(nsb_ctx **ctxs, int n);
int nsb_mgpu_group_step(nsb_ctx **ctxs, int n, double dt);
/* single-process group: bring ctxs[0]'s reference-layout OMEGA_HAT / SFD_HAT up to date from all
 * devices (gather) or publish it to all devices after the host changed it (scatter) */
int nsb_mgpu_group_gather(nsb_ctx **ctxs, int n);
int nsb_mgpu_group_scatter(nsb_ctx **ctxs, int n);

/* Turbulence statistics of a velocity field (/root/reference/src/analysis/statistics.c) on the device: the raw sums
 *   raw = { sum u, sum v, sum u'^2, sum u'^3, sum u'^4, sum v'^2, sum v'^3, sum v'^4,
 *           sum (du'/dx)^2 (centred, periodic), sum (Sxx^2 + Syy^2 + 2 Sxy^2) }
 * and the unnormalised two-point sums corr_x_raw[s] = sum u'(x) u'(x + s dx), s < sep_x (same along y); the
 * caller turns them into moments, correlation functions and length scales (csrc/host/analysis.c).
 * nsb_velocity_statistics works on the context's device-resident U, V (no transfer);
 * nsb_velocity_statistics_host takes host arrays of ANY size nx x ny (uploads them). Synchronous. */
int nsb_velocity_statistics(nsb_ctx *c, double dx, double dy, int sep_x, int sep_y, double raw[10],
                            double *corr_x_raw, double *corr_y_raw);
int nsb_velocity_statistics_host(int device, const double *u, const double *v, int nx, int ny, double dx, double dy,
                                 int sep_x, int sep_y, double raw[10], double *corr_x_raw, double *corr_y_raw);

/* Order-independent, exactly combinable fingerprint of the RK state omega_hat over the retained
 * columns this device owns ([col0, col1), all ky rows): out[0] = XOR, out[1] = wrapping 64-bit sum of
 * the IEEE-754 bit patterns of every real and imaginary part. XOR / add the results of all ranks of a
 * slab-decomposed run and compare with the one-device run: equal <=> the fields agree bit for bit
 * (up to XOR/sum collisions). Synchronous. */
int nsb_state_checksum(nsb_ctx *c, unsigned long long out[2]);

/* number of kernels this library has launched on this context */
long long nsb_launch_count(nsb_ctx *c);

#ifdef __cplusplus
}
#endif
#endif
