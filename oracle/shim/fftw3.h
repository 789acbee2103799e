/*
 * oracle/shim/fftw3.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Minimal stand-in for <fftw3.h> so the UNMODIFIED reference spectral sources
 * (/root/reference/src/spectral/ns_spectral_solver.h:6 includes <fftw3.h>)
 * compile in an image without FFTW3. Declares exactly the symbols the
 * reference uses (spectral_memory.c:24-44,74-102; spectral_initialization.c:58-68;
 * spectral_transforms.c:10,31). Implemented by fftw3_shim.c over oracle/fft_cpu.c.
 * Power-of-two sizes only: plan creation returns NULL otherwise.
 */
#ifndef NS_ORACLE_FFTW3_SHIM_H
#define NS_ORACLE_FFTW3_SHIM_H

#include <complex.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double complex fftw_complex;
typedef struct ns_shim_fftw_plan_s *fftw_plan;

#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

double *fftw_alloc_real(size_t n);
fftw_complex *fftw_alloc_complex(size_t n);
void fftw_free(void *p);

fftw_plan fftw_plan_dft_r2c_2d(int n0, int n1, double *in, fftw_complex *out, unsigned flags);
fftw_plan fftw_plan_dft_c2r_2d(int n0, int n1, fftw_complex *in, double *out, unsigned flags);
void fftw_execute_dft_r2c(const fftw_plan p, double *in, fftw_complex *out);
void fftw_execute_dft_c2r(const fftw_plan p, fftw_complex *in, double *out);
void fftw_destroy_plan(fftw_plan p);
void fftw_cleanup(void);

int fftw_init_threads(void);
void fftw_plan_with_nthreads(int nthreads);
void fftw_cleanup_threads(void);

#ifdef __cplusplus
}
#endif
#endif
