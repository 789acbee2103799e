/*
 * oracle/fft_cpu.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A small, accurate FP64 CPU FFT used by
 *   (1) oracle/shim/fftw3_shim.c : stands in for FFTW3 (not installed in this
 *       image, un-vendored by the reference: /root/reference/Makefile:10-12)
 *       so the UNMODIFIED reference sources can be built into oracle/_ref/;
 *   (2) oracle/ns_oracle.c       : the C restatement of the reference path.
 *
 * Conventions follow FFTW's, which is what the reference relies on
 * (/root/reference/src/spectral/spectral_transforms.c:10,31):
 *   - both directions unnormalised;
 *   - r2c output is ny x (nx/2+1), row-major, x fastest;
 *   - c2r destroys its input, and ignores Im of the DC and Nyquist bins of the
 *     last dimension after the complex pass over the first dimension.
 * Power-of-two sizes only (n >= 2).
 */
#ifndef NS_ORACLE_FFT_CPU_H
#define NS_ORACLE_FFT_CPU_H

#include <complex.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fftc_plan fftc_plan;

/* 1-D complex plan of length n (power of two). NULL on bad n / OOM. */
fftc_plan *fftc_plan_create(int n);
void fftc_plan_destroy(fftc_plan *p);
/* In-place transform of x[0..n); work[0..n) is scratch. sign=-1 forward, +1 inverse. */
void fftc_exec(const fftc_plan *p, double complex *x, double complex *work, int sign);

typedef struct fft2d_plan fft2d_plan;
fft2d_plan *fft2d_plan_create(int ny, int nx);
void fft2d_plan_destroy(fft2d_plan *p);
/* real ny x nx  ->  complex ny x (nx/2+1), unnormalised forward (e^{-i...}) */
void fft2d_r2c(const fft2d_plan *p, const double *in, double complex *out);
/* complex ny x (nx/2+1) -> real ny x nx, unnormalised inverse; `in` is overwritten */
void fft2d_c2r(const fft2d_plan *p, double complex *in, double *out);

#ifdef __cplusplus
}
#endif
#endif
