// Pairwise tile kernels (Striped-UniFrac style): one CTA owns one 128x128 tile of sample
// pairs and streams node stripes of the two sample blocks through shared memory.
//
// Replaces the reference's Python pair loop: scipy pdist calling _unweighted_unifrac /
// _weighted_unifrac[_normalized] once per pair
// (/root/reference/skbio/diversity/_driver.py:349-354,
//  /root/reference/skbio/diversity/beta/_unifrac.py:340-471), and scipy's compiled
// pdist 'braycurtis' / 'jaccard' loops for the feature-table metrics.
#pragma once

#include "b2d_common.cuh"

namespace b2d {

// Epilogue modes (applied when the last node range of a pass has been accumulated)
enum Epilogue : int {
    EPI_RAW = 0,            // weighted UniFrac (unnormalized): sum len*|pu-pv|
    EPI_WEIGHTED_NORM = 1,  // / (C_i + C_j); both totals zero -> 0   (_unifrac.py:461-471)
    EPI_UNWEIGHTED = 2,     // x / ((L_i + L_j + x)/2); denominator 0 -> 0 (_unifrac.py:366-371)
    EPI_BRAYCURTIS = 3,     // x / (S_i + S_j); 0/0 -> NaN like scipy
    EPI_JACCARD = 4         // x / ((c_i + c_j + x)/2); denominator 0 -> 0 like scipy
};

struct PairArgs {
    const int32_t* tiles;     // [ntiles][2] (row block, col block), col >= row
    const int64_t* row_base;  // [nb] offset of each owned row stripe in `out`
    double* out;              // compact shard result
    const double* svec;       // per-sample vector used by the epilogue (C, L, S or c)
    const long long* total;   // per-sample totals (weighted normalized only)
    int64_t n;                // samples
    int64_t qa, qb;           // node positions [qa, qb) relative to the resident range
    int64_t QR;               // node rows per block in the resident fp64 buffer
    int64_t NS;               // node stripes per block in the bits buffer
    int first, last;          // first: store, else accumulate; last: apply the epilogue
    int epilogue;
};

__device__ __forceinline__ double apply_epilogue(double x, int mode, const double* svec,
                                                 const long long* total, int64_t i, int64_t j) {
    switch (mode) {
        case EPI_WEIGHTED_NORM: {
            if (total[i] == 0 && total[j] == 0) return 0.0;
            return x / (svec[i] + svec[j]);
        }
        case EPI_UNWEIGHTED: {
            double den = 0.5 * (svec[i] + svec[j] + x);
            return den == 0.0 ? 0.0 : x / den;
        }
        case EPI_BRAYCURTIS:
            return x / (svec[i] + svec[j]);
        case EPI_JACCARD: {
            double den = 0.5 * (svec[i] + svec[j] + x);
            return den == 0.0 ? 0.0 : x / den;
        }
        default:
            return x;
    }
}

// Write one accumulated value of pair (i, j) (global sample indices).
__device__ __forceinline__ void emit(const PairArgs& a, int64_t base, int64_t i0, int64_t i,
                                     int64_t j, double v) {
    if (i >= j || j >= a.n) return;
    int64_t off = base + (cond_index(i, j, a.n) - cond_row_start(i0, a.n));
    double x = a.first ? v : a.out[off] + v;
    if (a.last) x = apply_epilogue(x, a.epilogue, a.svec, a.total, i, j);
    a.out[off] = x;
}

// =========================================================================================
// Weighted kernel: acc[i][j] += len_q * |p_q,i - p_q,j|   (also Bray-Curtis with len == 1)
// 256 compute threads (16x16, each an 8x8 register tile of fp64 accumulators) + one
// producer warp that feeds a 4-stage ring of {A rows, B rows, lengths} with bulk copies.
// FP64-pipe bound: 2 FP64 instructions (DADD + DFMA) per node-pair term.
// =========================================================================================
constexpr int WK_KS = 16;      // nodes per stage
constexpr int WK_STAGES = 4;
constexpr int WK_THREADS = 256 + 32;
constexpr int WK_STAGE_DOUBLES = 2 * WK_KS * TILE + WK_KS;  // A + B + len
constexpr size_t WK_SMEM = (size_t)WK_STAGES * WK_STAGE_DOUBLES * sizeof(double) + 2 * WK_STAGES * sizeof(uint64_t);

__global__ void __launch_bounds__(WK_THREADS, 1)
k_pair_weighted(const double* __restrict__ emb, const double* __restrict__ len_q, PairArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_base = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)WK_STAGES * WK_STAGE_DOUBLES * sizeof(double));
    uint64_t* empty = full + WK_STAGES;

    const int tid = threadIdx.x;
    const int bi = a.tiles[2 * blockIdx.x], bj = a.tiles[2 * blockIdx.x + 1];
    const int64_t niter = (a.qb - a.qa) / WK_KS;

    if (tid == 0) {
        for (int s = 0; s < WK_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 8);  // one arrive per compute warp
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (tid >= 256) {
        // ---------------- producer warp ----------------
        if (tid == 256) {
            const double* Ag = emb + ((int64_t)bi * a.QR + a.qa) * TILE;
            const double* Bg = emb + ((int64_t)bj * a.QR + a.qa) * TILE;
            const double* Lg = len_q + a.qa;
            for (int64_t it = 0; it < niter; ++it) {
                int s = (int)(it % WK_STAGES);
                uint32_t ph = (uint32_t)((it / WK_STAGES) & 1);
                if (it >= WK_STAGES) mbar_wait(&empty[s], ph ^ 1u);
                double* A = stage_base + (size_t)s * WK_STAGE_DOUBLES;
                double* B = A + WK_KS * TILE;
                double* L = B + WK_KS * TILE;
                mbar_arrive_expect_tx(&full[s], (uint32_t)(WK_STAGE_DOUBLES * sizeof(double)));
                bulk_g2s(A, Ag + it * WK_KS * TILE, WK_KS * TILE * sizeof(double), &full[s]);
                bulk_g2s(B, Bg + it * WK_KS * TILE, WK_KS * TILE * sizeof(double), &full[s]);
                bulk_g2s(L, Lg + it * WK_KS, WK_KS * sizeof(double), &full[s]);
            }
        }
        return;
    }

    // ---------------- compute warps ----------------
    const int tx = tid & 15, ty = tid >> 4;
    double acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.0;

    for (int64_t it = 0; it < niter; ++it) {
        int s = (int)(it % WK_STAGES);
        uint32_t ph = (uint32_t)((it / WK_STAGES) & 1);
        mbar_wait(&full[s], ph);
        const double* A = stage_base + (size_t)s * WK_STAGE_DOUBLES;
        const double* B = A + WK_KS * TILE;
        const double* L = B + WK_KS * TILE;
#pragma unroll 2
        for (int k = 0; k < WK_KS; ++k) {
            double av[8], bv[8];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                double2 t = *reinterpret_cast<const double2*>(A + k * TILE + h * 32 + ty * 2);
                av[2 * h] = t.x;
                av[2 * h + 1] = t.y;
                double2 u = *reinterpret_cast<const double2*>(B + k * TILE + h * 32 + tx * 2);
                bv[2 * h] = u.x;
                bv[2 * h + 1] = u.y;
            }
            const double l = L[k];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = fma(fabs(av[r] - bv[c]), l, acc[r][c]);
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    }

    // ---------------- epilogue ----------------
    const int64_t base = a.row_base[bi];
    const int64_t i0 = (int64_t)bi * TILE;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int64_t i = i0 + (r >> 1) * 32 + ty * 2 + (r & 1);
        if (i >= a.n) continue;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            int64_t j = (int64_t)bj * TILE + (c >> 1) * 32 + tx * 2 + (c & 1);
            emit(a, base, i0, i, j, acc[r][c]);
        }
    }
}

// Plain variant of the same arithmetic (no bulk copies, no mbarriers): cross-check for tests
// and a fallback when B2D_PLAIN_KERNELS=1.
__global__ void __launch_bounds__(256, 1)
k_pair_weighted_plain(const double* __restrict__ emb, const double* __restrict__ len_q, PairArgs a) {
    __shared__ double A[WK_KS * TILE];
    __shared__ double B[WK_KS * TILE];
    __shared__ double L[WK_KS];
    const int tid = threadIdx.x;
    const int bi = a.tiles[2 * blockIdx.x], bj = a.tiles[2 * blockIdx.x + 1];
    const int tx = tid & 15, ty = tid >> 4;
    double acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.0;
    const double* Ag = emb + ((int64_t)bi * a.QR + a.qa) * TILE;
    const double* Bg = emb + ((int64_t)bj * a.QR + a.qa) * TILE;
    for (int64_t q = 0; q < a.qb - a.qa; q += WK_KS) {
        __syncthreads();
        for (int x = tid; x < WK_KS * TILE; x += 256) {
            A[x] = Ag[q * TILE + x];
            B[x] = Bg[q * TILE + x];
        }
        if (tid < WK_KS) L[tid] = len_q[a.qa + q + tid];
        __syncthreads();
        for (int k = 0; k < WK_KS; ++k) {
            double av[8], bv[8];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                av[2 * h] = A[k * TILE + h * 32 + ty * 2];
                av[2 * h + 1] = A[k * TILE + h * 32 + ty * 2 + 1];
                bv[2 * h] = B[k * TILE + h * 32 + tx * 2];
                bv[2 * h + 1] = B[k * TILE + h * 32 + tx * 2 + 1];
            }
            const double l = L[k];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = fma(fabs(av[r] - bv[c]), l, acc[r][c]);
        }
    }
    const int64_t base = a.row_base[bi];
    const int64_t i0 = (int64_t)bi * TILE;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int64_t i = i0 + (r >> 1) * 32 + ty * 2 + (r & 1);
        if (i >= a.n) continue;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            int64_t j = (int64_t)bj * TILE + (c >> 1) * 32 + tx * 2 + (c & 1);
            emit(a, base, i0, i, j, acc[r][c]);
        }
    }
}

// =========================================================================================
// Unweighted kernel (formulation S on node-packed bits): acc[i][j] += len_q * [u_q ^ v_q].
// Per 128-node stripe a CTA stages 2 KB of row bits, 2 KB of column bits and 1 KB of branch
// lengths; each thread owns an 8x8 register tile of fp64 accumulators and issues one
// predicated DADD per set xor bit.
// =========================================================================================
constexpr int UK_STAGES = 2;
struct __align__(16) UStage {
    uint4 rows[TILE];
    uint4 cols[TILE];
    double len[NODE_STRIPE];
};

__global__ void __launch_bounds__(256, 1)
k_pair_unweighted(const uint4* __restrict__ bits, const double* __restrict__ len_q, PairArgs a) {
    __shared__ UStage st[UK_STAGES];
    const int tid = threadIdx.x;
    const int bi = a.tiles[2 * blockIdx.x], bj = a.tiles[2 * blockIdx.x + 1];
    const int tx = tid & 15, ty = tid >> 4;
    double acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.0;

    const int64_t s0 = a.qa / NODE_STRIPE, s1 = a.qb / NODE_STRIPE;
    const uint4* Rg = bits + ((int64_t)bi * a.NS) * TILE;
    const uint4* Cg = bits + ((int64_t)bj * a.NS) * TILE;

    auto load_stage = [&](int buf, int64_t s) {
        if (tid < TILE)
            st[buf].rows[tid] = Rg[s * TILE + tid];
        else
            st[buf].cols[tid - TILE] = Cg[s * TILE + (tid - TILE)];
        if (tid < NODE_STRIPE) st[buf].len[tid] = len_q[s * NODE_STRIPE + tid];
    };

    if (s0 < s1) load_stage(0, s0);
    for (int64_t s = s0; s < s1; ++s) {
        const int buf = (int)((s - s0) & 1);
        __syncthreads();
        if (s + 1 < s1) load_stage(buf ^ 1, s + 1);
        const uint32_t* R = reinterpret_cast<const uint32_t*>(st[buf].rows);
        const uint32_t* C = reinterpret_cast<const uint32_t*>(st[buf].cols);
#pragma unroll 1
        for (int w = 0; w < 4; ++w) {
            uint32_t rw[8], cw[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                rw[e] = R[(ty * 8 + e) * 4 + w];
                cw[e] = C[(e * 16 + tx) * 4 + w];
            }
#pragma unroll 1
            for (int by = 0; by < 4; ++by) {
                double l[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) l[t] = st[buf].len[w * 32 + by * 8 + t];
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        uint32_t x = ((rw[r] ^ cw[c]) >> (8 * by)) & 0xffu;
#pragma unroll
                        for (int t = 0; t < 8; ++t)
                            if (x & (1u << t)) acc[r][c] += l[t];
                    }
            }
        }
    }

    const int64_t base = a.row_base[bi];
    const int64_t i0 = (int64_t)bi * TILE;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int64_t i = i0 + ty * 8 + r;
        if (i >= a.n) continue;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            int64_t j = (int64_t)bj * TILE + c * 16 + tx;
            emit(a, base, i0, i, j, acc[r][c]);
        }
    }
}

// =========================================================================================
// Jaccard kernel: acc[i][j] += popcount(u ^ v) over feature stripes (exact integers).
// =========================================================================================
__global__ void __launch_bounds__(256, 2)
k_pair_popcount(const uint4* __restrict__ bits, PairArgs a) {
    __shared__ uint4 rows[2][TILE];
    __shared__ uint4 cols[2][TILE];
    const int tid = threadIdx.x;
    const int bi = a.tiles[2 * blockIdx.x], bj = a.tiles[2 * blockIdx.x + 1];
    const int tx = tid & 15, ty = tid >> 4;
    unsigned acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0u;
    const int64_t s0 = a.qa / NODE_STRIPE, s1 = a.qb / NODE_STRIPE;
    const uint4* Rg = bits + ((int64_t)bi * a.NS) * TILE;
    const uint4* Cg = bits + ((int64_t)bj * a.NS) * TILE;
    auto load_stage = [&](int buf, int64_t s) {
        if (tid < TILE)
            rows[buf][tid] = Rg[s * TILE + tid];
        else
            cols[buf][tid - TILE] = Cg[s * TILE + (tid - TILE)];
    };
    if (s0 < s1) load_stage(0, s0);
    for (int64_t s = s0; s < s1; ++s) {
        const int buf = (int)((s - s0) & 1);
        __syncthreads();
        if (s + 1 < s1) load_stage(buf ^ 1, s + 1);
        const uint32_t* R = reinterpret_cast<const uint32_t*>(rows[buf]);
        const uint32_t* C = reinterpret_cast<const uint32_t*>(cols[buf]);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint32_t rw[8], cw[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                rw[e] = R[(ty * 8 + e) * 4 + w];
                cw[e] = C[(e * 16 + tx) * 4 + w];
            }
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] += __popc(rw[r] ^ cw[c]);
        }
    }
    const int64_t base = a.row_base[bi];
    const int64_t i0 = (int64_t)bi * TILE;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int64_t i = i0 + ty * 8 + r;
        if (i >= a.n) continue;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            int64_t j = (int64_t)bj * TILE + c * 16 + tx;
            emit(a, base, i0, i, j, (double)acc[r][c]);
        }
    }
}

}  // namespace b2d
