/*
 * oracle/fft_cpu.c -- TEST INFRASTRUCTURE ONLY. See fft_cpu.h.
 *
 * Stockham autosort FFT, radix-4 with a trailing radix-2 level, FP64, twiddles
 * tabulated once in long double. Inverse = conj(forward(conj(x))). The 2-D
 * real transforms use the half-length complex trick along x and blocked
 * gathers along y; OpenMP over rows / column blocks.
 */
#include "fft_cpu.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

struct fftc_plan {
    int n;
    double complex *w; /* w[k] = exp(-2*pi*i*k/n), k in [0,n) */
};

static int is_pow2(int n) { return n >= 1 && (n & (n - 1)) == 0; }

fftc_plan *fftc_plan_create(int n) {
    if (!is_pow2(n)) return NULL;
    fftc_plan *p = (fftc_plan *)malloc(sizeof(*p));
    if (!p) return NULL;
    p->n = n;
    p->w = (double complex *)malloc(sizeof(double complex) * (size_t)n);
    if (!p->w) { free(p); return NULL; }
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int k = 0; k < n; ++k) {
        /* reduce to the first octant so cosl/sinl see a small argument */
        int kk = k;
        int quad = (int)((4LL * kk) / n);       /* 0..3 */
        int rem = kk - (int)(((long long)quad * n) / 4);
        long double a = two_pi * (long double)rem / (long double)n;
        long double c = cosl(a), s = sinl(a);
        long double cr, ci;                     /* exp(-i*(quad*pi/2 + a)) */
        switch (quad) {
            case 0: cr = c;  ci = -s; break;
            case 1: cr = -s; ci = -c; break;
            case 2: cr = -c; ci = s;  break;
            default: cr = s; ci = c;  break;
        }
        p->w[k] = (double)cr + (double)ci * I;
    }
    return p;
}

void fftc_plan_destroy(fftc_plan *p) {
    if (!p) return;
    free(p->w);
    free(p);
}

/* forward transform, result ends in x */
static void stockham_forward(const fftc_plan *p, double complex *x, double complex *y) {
    const int N = p->n;
    const double complex *w = p->w;
    int n = N, s = 1;
    double complex *src = x, *dst = y;
    while (n >= 4) {
        const int n1 = n / 4, n2 = n / 2, n3 = n1 + n2;
        for (int q0 = 0; q0 < n1; ++q0) {
            const double complex w1 = w[(size_t)q0 * s];
            const double complex w2 = w[(size_t)2 * q0 * s];
            const double complex w3 = w[(size_t)3 * q0 * s];
            const double complex *a = src + (size_t)s * q0;
            const double complex *b = src + (size_t)s * (q0 + n1);
            const double complex *c = src + (size_t)s * (q0 + n2);
            const double complex *d = src + (size_t)s * (q0 + n3);
            double complex *o = dst + (size_t)s * 4 * q0;
            for (int q = 0; q < s; ++q) {
                const double complex apc = a[q] + c[q], amc = a[q] - c[q];
                const double complex bpd = b[q] + d[q];
                const double complex bmd = b[q] - d[q];
                /* i*(b-d) */
                const double complex jbmd = -cimag(bmd) + creal(bmd) * I;
                o[q] = apc + bpd;
                o[q + s] = w1 * (amc - jbmd);
                o[q + 2 * s] = w2 * (apc - bpd);
                o[q + 3 * s] = w3 * (amc + jbmd);
            }
        }
        double complex *t = src; src = dst; dst = t;
        n /= 4; s *= 4;
    }
    if (n == 2) {
        const double complex *a = src, *b = src + s;
        for (int q = 0; q < s; ++q) {
            const double complex u = a[q], v = b[q];
            dst[q] = u + v;
            dst[q + s] = u - v;
        }
        double complex *t = src; src = dst; dst = t;
    }
    if (src != x) memcpy(x, src, sizeof(double complex) * (size_t)N);
}

void fftc_exec(const fftc_plan *p, double complex *x, double complex *work, int sign) {
    if (p->n == 1) return;
    if (sign < 0) {
        stockham_forward(p, x, work);
    } else {
        for (int k = 0; k < p->n; ++k) x[k] = conj(x[k]);
        stockham_forward(p, x, work);
        for (int k = 0; k < p->n; ++k) x[k] = conj(x[k]);
    }
}

/* ------------------------------------------------------------------ 2-D */

struct fft2d_plan {
    int ny, nx, nkx;
    fftc_plan *px_half;       /* length nx/2 */
    fftc_plan *py;            /* length ny   */
    double complex *wx;       /* exp(-2*pi*i*k/nx), k in [0, nx/2] */
};

#define COLBLK 8

fft2d_plan *fft2d_plan_create(int ny, int nx) {
    if (!is_pow2(nx) || !is_pow2(ny) || nx < 2 || ny < 1) return NULL;
    fft2d_plan *p = (fft2d_plan *)calloc(1, sizeof(*p));
    if (!p) return NULL;
    p->ny = ny; p->nx = nx; p->nkx = nx / 2 + 1;
    p->px_half = fftc_plan_create(nx / 2);
    p->py = fftc_plan_create(ny);
    fftc_plan *full = fftc_plan_create(nx);
    p->wx = (double complex *)malloc(sizeof(double complex) * (size_t)(nx / 2 + 1));
    if (!p->px_half || !p->py || !full || !p->wx) {
        fftc_plan_destroy(full);
        fft2d_plan_destroy(p);
        return NULL;
    }
    for (int k = 0; k < nx / 2; ++k) p->wx[k] = full->w[k];
    p->wx[nx / 2] = -1.0;
    fftc_plan_destroy(full);
    return p;
}

void fft2d_plan_destroy(fft2d_plan *p) {
    if (!p) return;
    fftc_plan_destroy(p->px_half);
    fftc_plan_destroy(p->py);
    free(p->wx);
    free(p);
}

static void columns_pass(const fft2d_plan *p, double complex *a, int sign) {
    const int ny = p->ny, nkx = p->nkx;
    if (ny == 1) return;
    const int nblk = (nkx + COLBLK - 1) / COLBLK;
#pragma omp parallel
    {
        double complex *buf = (double complex *)malloc(sizeof(double complex) * (size_t)ny * (COLBLK + 1));
        double complex *work = buf + (size_t)ny * COLBLK;
#pragma omp for schedule(static)
        for (int b = 0; b < nblk; ++b) {
            const int i0 = b * COLBLK;
            const int nc = (i0 + COLBLK <= nkx) ? COLBLK : nkx - i0;
            for (int j = 0; j < ny; ++j)
                for (int c = 0; c < nc; ++c)
                    buf[(size_t)c * ny + j] = a[(size_t)j * nkx + i0 + c];
            for (int c = 0; c < nc; ++c)
                fftc_exec(p->py, buf + (size_t)c * ny, work, sign);
            for (int j = 0; j < ny; ++j)
                for (int c = 0; c < nc; ++c)
                    a[(size_t)j * nkx + i0 + c] = buf[(size_t)c * ny + j];
        }
        free(buf);
    }
}

void fft2d_r2c(const fft2d_plan *p, const double *in, double complex *out) {
    const int ny = p->ny, nx = p->nx, nkx = p->nkx, h = nx / 2;
#pragma omp parallel
    {
        double complex *z = (double complex *)malloc(sizeof(double complex) * (size_t)h * 2);
        double complex *work = z + h;
#pragma omp for schedule(static)
        for (int j = 0; j < ny; ++j) {
            const double *row = in + (size_t)j * nx;
            double complex *o = out + (size_t)j * nkx;
            for (int m = 0; m < h; ++m) z[m] = row[2 * m] + row[2 * m + 1] * I;
            fftc_exec(p->px_half, z, work, -1);
            for (int k = 0; k <= h; ++k) {
                const double complex zk = z[k % h];
                const double complex zc = conj(z[(h - k) % h]);
                const double complex e = 0.5 * (zk + zc);
                const double complex d = 0.5 * (zk - zc);
                /* X[k] = E[k] - i*w^k*D[k] */
                const double complex wd = p->wx[k] * d;
                o[k] = e + (cimag(wd) - creal(wd) * I);
            }
        }
        free(z);
    }
    columns_pass(p, out, -1);
}

void fft2d_c2r(const fft2d_plan *p, double complex *in, double *out) {
    const int ny = p->ny, nx = p->nx, nkx = p->nkx, h = nx / 2;
    columns_pass(p, in, +1);
#pragma omp parallel
    {
        double complex *z = (double complex *)malloc(sizeof(double complex) * (size_t)h * 2);
        double complex *work = z + h;
#pragma omp for schedule(static)
        for (int j = 0; j < ny; ++j) {
            const double complex *X = in + (size_t)j * nkx;
            double *row = out + (size_t)j * nx;
            for (int k = 0; k < h; ++k) {
                double complex xk = X[k];
                double complex xm = X[h - k];
                if (k == 0) { xk = creal(xk); xm = creal(xm); } /* Im(DC), Im(Nyquist) ignored */
                const double complex e = xk + conj(xm);
                const double complex d = xk - conj(xm);
                /* Z[k] = E + i*conj(w^k)*D */
                const double complex wd = conj(p->wx[k]) * d;
                z[k] = e + (-cimag(wd) + creal(wd) * I);
            }
            fftc_exec(p->px_half, z, work, +1);
            for (int m = 0; m < h; ++m) {
                row[2 * m] = creal(z[m]);
                row[2 * m + 1] = cimag(z[m]);
            }
        }
        free(z);
    }
}
