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

#endif
