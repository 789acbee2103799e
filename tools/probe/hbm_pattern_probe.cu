// Development probe (not part of the product): how fast does HBM3e take the access patterns of
// the step's kernels? Each test moves `bytes` in segments of `seg` bytes whose starts are
// `stride` bytes apart within a "row group", the way the column kernel (64/128-byte segments,
// one per T row, rows 16*PT*16 bytes apart), B (256-byte segments of 16 rows PT*16 bytes apart)
// and the Z writes (256-byte segments 16 KB apart) do. Build: nvcc -arch=sm_100a -O3 -o probe ...
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

// One warp moves one segment per iteration; consecutive warps of the grid take consecutive
// segments of the SAME row (so neighbouring CTAs touch neighbouring bytes, like the real kernels),
// then the next row `stride` bytes further.
template <bool WRITE, bool READ>
__global__ void k_pattern(double2 *buf, size_t nrows, size_t row_bytes, size_t stride, int seg) {
    const int lanes_per_seg = seg / 16;                 // 16-byte elements
    const int segs_per_warp = 32 / lanes_per_seg;
    const size_t segs_per_row = row_bytes / seg;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) / 32;
    const int lane = threadIdx.x & 31;
    const int sub = lane / lanes_per_seg, l = lane % lanes_per_seg;
    double2 acc = make_double2(0.0, 0.0);
    // work item = (column group, block of 1024 rows); consecutive warps take consecutive column
    // groups of the same row block, like consecutive CTAs of the column kernel
    const size_t ncg = (segs_per_row + segs_per_warp - 1) / segs_per_warp;
    const size_t nrb = nrows / 1024;
    for (size_t item = warp; item < ncg * nrb; item += nwarps) {
        const size_t rb = item / ncg, cg = item - rb * ncg;
        const size_t col = cg * segs_per_warp + sub;
        if (col >= segs_per_row) continue;
#pragma unroll 8
        for (size_t r = rb * 1024; r < (rb + 1) * 1024; ++r) {
            double2 *p = (double2 *)((char *)buf + r * stride + col * seg) + l;
            if (READ) { const double2 v = *p; acc.x += v.x; acc.y += v.y; }
            if (WRITE) *p = make_double2((double)r, (double)col);
        }
    }
    if (READ && acc.x == 1.2345e300) buf[0] = acc;
}

template <bool W, bool R>
static void run(const char *name, double2 *buf, size_t nrows, size_t row_bytes, size_t stride, int seg) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256;
    k_pattern<W, R><<<grid, block>>>(buf, nrows, row_bytes, stride, seg);
    cudaEventRecord(e0);
    const int reps = 3;
    for (int i = 0; i < reps; ++i) k_pattern<W, R><<<grid, block>>>(buf, nrows, row_bytes, stride, seg);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)nrows * row_bytes * ((W ? 1 : 0) + (R ? 1 : 0));
    printf("%-58s seg=%4d B stride=%8zu B : %7.0f GB/s  (%s)\n", name, seg, stride, bytes * reps / (ms * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t row_bytes = 87424;            // PT*16 at 16384^2
    const size_t nrows = 65536;                // 4 fields x 16384 rows: 5.7 GB
    double2 *buf;
    if (cudaMalloc(&buf, nrows * row_bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 0, nrows * row_bytes);
    // rows contiguous (stride == row): a warp column walks rows 87 KB apart -> "one segment per row"
    run<true, false>("write, one segment per row (column kernel, 4 cols)", buf, nrows, row_bytes, row_bytes, 64);
    run<true, false>("write, one segment per row (column kernel, 8 cols)", buf, nrows, row_bytes, row_bytes, 128);
    run<true, false>("write, one segment per row (16 cols)", buf, nrows, row_bytes, row_bytes, 256);
    run<true, false>("write, one segment per row (32 cols)", buf, nrows, row_bytes, row_bytes, 512);
    run<false, true>("read, one segment per row (B boxes, 16 cols)", buf, nrows, row_bytes, row_bytes, 256);
    run<false, true>("read, one segment per row (8 cols)", buf, nrows, row_bytes, row_bytes, 128);
    run<false, true>("read, one segment per row (4 cols)", buf, nrows, row_bytes, row_bytes, 64);
    // column-kernel-like for real: consecutive rows of a tile are 16 T rows (1.4 MB) apart, so nearly
    // every row of a 1024-row tile lies in a different 2 MB page (TLB reach: 128 entries)
    run<true, false>("write, tile rows 16 T-rows apart (alpha, 8 cols)", buf, nrows / 16, row_bytes, 16 * row_bytes, 128);
    run<true, false>("write, tile rows 16 T-rows apart (alpha, 4 cols)", buf, nrows / 16, row_bytes, 16 * row_bytes, 64);
    run<false, true>("read, tile rows 16 T-rows apart (alpha T' load, 8 cols)", buf, nrows / 16, row_bytes, 16 * row_bytes, 128);
    // Z-like: rows of 16 KB
    run<true, false>("write, 256 B per 16 KB row (Z)", buf, (nrows * row_bytes) / 16384, 16384, 16384, 256);
    run<true, true>("read+write same pattern 128 B", buf, nrows, row_bytes, row_bytes, 128);
    cudaFree(buf);
    return 0;
}
