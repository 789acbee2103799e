/*
 * oracle/shim/fftw3_shim.c -- TEST INFRASTRUCTURE ONLY. See fftw3.h (shim).
 * Maps the 13 FFTW symbols the reference uses onto oracle/fft_cpu.c.
 */
#include "fftw3.h"
#include "../fft_cpu.h"

#include <stdlib.h>

struct ns_shim_fftw_plan_s {
    fft2d_plan *p2d;
};

double *fftw_alloc_real(size_t n) {
    void *p = NULL;
    if (posix_memalign(&p, 64, (n ? n : 1) * sizeof(double)) != 0) return NULL;
    return (double *)p;
}

fftw_complex *fftw_alloc_complex(size_t n) {
    void *p = NULL;
    if (posix_memalign(&p, 64, (n ? n : 1) * sizeof(fftw_complex)) != 0) return NULL;
    return (fftw_complex *)p;
}

void fftw_free(void *p) { free(p); }

static fftw_plan make_plan(int n0, int n1) {
    fftw_plan pl = (fftw_plan)calloc(1, sizeof(*pl));
    if (!pl) return NULL;
    pl->p2d = fft2d_plan_create(n0, n1);
    if (!pl->p2d) { free(pl); return NULL; }
    return pl;
}

fftw_plan fftw_plan_dft_r2c_2d(int n0, int n1, double *in, fftw_complex *out, unsigned flags) {
    (void)in; (void)out; (void)flags;
    return make_plan(n0, n1);
}

fftw_plan fftw_plan_dft_c2r_2d(int n0, int n1, fftw_complex *in, double *out, unsigned flags) {
    (void)in; (void)out; (void)flags;
    return make_plan(n0, n1);
}

void fftw_execute_dft_r2c(const fftw_plan p, double *in, fftw_complex *out) {
    fft2d_r2c(p->p2d, in, out);
}

void fftw_execute_dft_c2r(const fftw_plan p, fftw_complex *in, double *out) {
    fft2d_c2r(p->p2d, in, out);
}

void fftw_destroy_plan(fftw_plan p) {
    if (!p) return;
    fft2d_plan_destroy(p->p2d);
    free(p);
}

void fftw_cleanup(void) {}
int fftw_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int nthreads) { (void)nthreads; }
void fftw_cleanup_threads(void) {}
