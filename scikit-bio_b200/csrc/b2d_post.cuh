// Consumers of the device-resident condensed result ("next" rows of SURVEY.md section 8f):
//   * k_square_rows: a block of rows of the square (redundant) matrix gathered from the condensed
//     vector - what a streaming lsmat / binary_dm writer needs row block by row block
//     (/root/reference/skbio/io/format/lsmat.py:266-279 writes obj.data row by row);
//   * Gower centering, the first step of pcoa
//     (/root/reference/skbio/stats/ordination/_cutils.pyx:19-123): E = -d^2/2 (e_matrix_means_cy),
//     row means and global mean, then F = E + ((mean - mean_row) - mean_col) (f_matrix_inplace_cy,
//     same association order).  Sums: per-thread sequential, then a fixed tree - deterministic,
//     ~1e-15 relative (the reference accumulates in long double).
#pragma once

#include "b2d_common.cuh"

namespace b2d {

__device__ __forceinline__ double cond_at(const double* __restrict__ cond, int64_t i, int64_t j, int64_t n) {
    return i == j ? 0.0 : (i < j ? cond[cond_index(i, j, n)] : cond[cond_index(j, i, n)]);
}

// out[(i - r0) * n + j] = d(i, j) for rows r0 <= i < r1; one CTA per row
__global__ void __launch_bounds__(256)
k_square_rows(const double* __restrict__ cond, int64_t n, int64_t r0, double* __restrict__ out) {
    const int64_t i = r0 + blockIdx.x;
    double* o = out + (int64_t)blockIdx.x * n;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) o[j] = cond_at(cond, i, j, n);
}

__device__ __forceinline__ double block_sum_fixed(double v, double* sh) {
    const int tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (tid < o) sh[tid] += sh[tid + o];
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

// row_sum[i] = sum_j -0.5 * d_ij * d_ij ; row_mean[i] = row_sum[i] / n
__global__ void __launch_bounds__(256)
k_center_rowsums(const double* __restrict__ cond, int64_t n, double* __restrict__ row_sum,
                 double* __restrict__ row_mean) {
    __shared__ double sh[256];
    const int64_t i = blockIdx.x;
    double s = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += 256) {
        const double d = cond_at(cond, i, j, n);
        s += -0.5 * d * d;
    }
    s = block_sum_fixed(s, sh);
    if (threadIdx.x == 0) {
        row_sum[i] = s;
        row_mean[i] = s / (double)n;
    }
}

// global mean = (sum of the row sums / n) / n, one CTA
__global__ void __launch_bounds__(256)
k_center_global(const double* __restrict__ row_sum, int64_t n, double* __restrict__ global_mean) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) s += row_sum[i];
    s = block_sum_fixed(s, sh);
    if (threadIdx.x == 0) *global_mean = (s / (double)n) / (double)n;
}

// rows r0.. of F: out[(i - r0) * n + j] = E_ij + ((gm - rm_i) - rm_j)
__global__ void __launch_bounds__(256)
k_center_fill(const double* __restrict__ cond, int64_t n, int64_t r0, const double* __restrict__ row_mean,
              const double* __restrict__ global_mean, double* __restrict__ out) {
    const int64_t i = r0 + blockIdx.x;
    const double gr = *global_mean - row_mean[i];
    double* o = out + (int64_t)blockIdx.x * n;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
        const double d = cond_at(cond, i, j, n);
        o[j] = -0.5 * d * d + (gr - row_mean[j]);
    }
}

}  // namespace b2d
