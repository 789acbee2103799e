/*
 * include/common_types.h -- constants and helper macros shared by the spectral
 * entry points. Mirrors the parts of /root/reference/include/common_types.h the
 * spectral path uses (NS_RESTRICT :9-17, M_PI :26-28, MIN_DT/MAX_DT :32-33,
 * SolverStatus :51-57). The finite-difference-only types are out of scope.
 */
#ifndef COMMON_TYPES_H
#define COMMON_TYPES_H

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if defined(__cplusplus)
#define NS_RESTRICT __restrict__
#elif defined(__STDC_VERSION__) && __STDC_VERSION__ >= 199901L
#define NS_RESTRICT restrict
#elif defined(__GNUC__)
#define NS_RESTRICT __restrict__
#else
#define NS_RESTRICT
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define MIN_DT 1e-8
#define MAX_DT 0.1

typedef enum {
    SOLVER_SUCCESS = 0,
    SOLVER_MAX_ITER,
    SOLVER_DIVERGED,
    SOLVER_MEM_ERROR,
    SOLVER_INVALID_PARAM
} SolverStatus;

#endif /* COMMON_TYPES_H */
