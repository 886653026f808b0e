// PERMANOVA pseudo-F partial statistic s_W for many groupings at once, on a condensed distance
// vector resident on the device.  Replaces the per-permutation calls of
// permanova_f_stat_sW[_condensed]_cy (/root/reference/skbio/stats/distance/_cutils.pyx:371-470)
// driven by _run_monte_carlo_stats (/root/reference/skbio/stats/distance/_base.py:2340-2364):
//     s_W(g) = sum_{i<j, g[i]==g[j]} d_ij^2 / |group of i|
// One CTA = (row stripe of 128 samples, batch of 16 groupings): it sweeps the column stripes,
// squares a 128x128 tile of distances once and tests it against the 16 groupings staged in
// shared memory; per-thread accumulators, fixed-order block reduction, one partial per
// (row stripe, grouping) - deterministic, no atomics.  s_T is obtained with one extra
// "everything in one group" row whose weight is 1/n.
#pragma once

#include "b2d_common.cuh"

namespace b2d {

constexpr int PM_BATCH = 16;

__global__ void __launch_bounds__(256)
k_permanova_sw(const double* __restrict__ cond, int64_t n, const int32_t* __restrict__ groupings,
               int64_t n_groupings, const double* __restrict__ inv_size, int n_groups,
               double* __restrict__ partial /*[nb][n_groupings]*/) {
    __shared__ int s_grow[PM_BATCH][TILE];
    __shared__ int s_gcol[PM_BATCH][TILE];
    __shared__ double s_red[8][PM_BATCH];
    extern __shared__ double s_w[];  // [n_groups]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t bi = blockIdx.x, nb = gridDim.x;
    const int64_t p0 = (int64_t)blockIdx.y * PM_BATCH;
    for (int g = tid; g < n_groups; g += 256) s_w[g] = inv_size[g];
    for (int x = tid; x < PM_BATCH * TILE; x += 256) {
        const int p = x / TILE, r = x % TILE;
        const int64_t i = bi * TILE + r;
        s_grow[p][r] = (p0 + p < n_groupings && i < n) ? groupings[(p0 + p) * n + i] : -1;
    }
    double acc[PM_BATCH];
#pragma unroll
    for (int p = 0; p < PM_BATCH; ++p) acc[p] = 0.0;

    for (int64_t bj = bi; bj < nb; ++bj) {
        __syncthreads();
        for (int x = tid; x < PM_BATCH * TILE; x += 256) {
            const int p = x / TILE, c = x % TILE;
            const int64_t j = bj * TILE + c;
            s_gcol[p][c] = (p0 + p < n_groupings && j < n) ? groupings[(p0 + p) * n + j] : -2;
        }
        __syncthreads();
        // this thread's 8 x 8 squared distances: rows ty*8+r, columns c*16+tx
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            double d2[4][8];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int64_t i = bi * TILE + ty * 8 + half * 4 + r;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int64_t j = bj * TILE + c * 16 + tx;
                    double v = 0.0;
                    if (i < j && j < n) v = cond[cond_index(i, j, n)];
                    d2[r][c] = v * v;
                }
            }
#pragma unroll
            for (int p = 0; p < PM_BATCH; ++p) {
                int gc[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) gc[c] = s_gcol[p][c * 16 + tx];
                double a = acc[p];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int gr = s_grow[p][ty * 8 + half * 4 + r];
                    const double w = gr >= 0 ? s_w[gr] : 0.0;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (gc[c] == gr) a = fma(d2[r][c], w, a);
                }
                acc[p] = a;
            }
        }
    }
    // fixed-order reduction: lanes (xor tree), then warps 0..7 in order
#pragma unroll
    for (int p = 0; p < PM_BATCH; ++p) {
        double v = acc[p];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) s_red[tid >> 5][p] = v;
    }
    __syncthreads();
    if (tid < PM_BATCH && p0 + tid < n_groupings) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += s_red[w][tid];
        partial[bi * n_groupings + p0 + tid] = v;
    }
}

__global__ void k_permanova_reduce(const double* __restrict__ partial, int64_t nb, int64_t n_groupings,
                                   double* __restrict__ out) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_groupings) return;
    double v = 0.0;
    for (int64_t b = 0; b < nb; ++b) v += partial[b * n_groupings + p];
    out[p] = v;
}

}  // namespace b2d
