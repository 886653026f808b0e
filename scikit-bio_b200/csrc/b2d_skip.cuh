// Zero-block skipping for the weighted / dense Bray-Curtis pair kernel.
//
// Samples are cut into groups of 32 (one warp patch of the pair kernel is 32 x 32 samples).
// For group g and node row q let z[g][q] = 1 when every sample of g is zero at q.  For a pair
// (i, j) a node with z[g(i)][q] = 1 contributes len_q * |0 - p_qj| = len_q * p_qj whatever i
// is, and a node with z[g(j)][q] = 1 contributes len_q * p_qi (0 when both are zero).  With
//   S[g][s] = sum_q z[g][q] * len_q * p_qs
//   x(i,j)  = sum_{q: z[g(i)][q] = 0 and z[g(j)][q] = 0} len_q |p_qi - p_qj|  +  S[g(i)][j] + S[g(j)][i]
// exactly (same non-negative terms, another summation order).  The pair kernel only visits the
// nodes of the first sum, decided per consumer warp (32 x 32 patch) from the bitmaps z.
#pragma once

#include "b2d_common.cuh"
#include "b2d_pair.cuh"

namespace b2d {

// nz[g][w] bit b = some sample of 32-sample group g is non-zero at resident node row 32w+b.
// One warp per (group, 32 node rows); lane = sample.
__global__ void __launch_bounds__(256)
k_group_nonzero(const double* __restrict__ emb, int64_t ngroups, int64_t rows, int64_t QR,
                uint32_t* __restrict__ nz /*[ngroups][QR/32]*/) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t words = QR / 32;
    if (warp >= ngroups * words) return;
    const int64_t g = warp / words, w = warp % words;
    const double* p = emb + ((g >> 2) * QR + w * 32) * TILE + (g & 3) * 32 + lane;
    uint32_t bits = 0;
#pragma unroll 8
    for (int b = 0; b < 32; ++b) {
        const bool any = (w * 32 + b < rows) && (p[(int64_t)b * TILE] != 0.0);
        if (__any_sync(0xffffffffu, any)) bits |= 1u << b;
    }
    if (lane == 0) nz[g * words + w] = bits;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// S[g][s] += sum over resident rows q with z[g][q] = 1 of len_q * p_qs.
// One warp = 32 groups (lane g: group glist[blockIdx.y*32 + g]) x the 32 samples of one sample
// group (4 warps of a CTA = the 4 sample groups of block sblist[blockIdx.x]); the warp walks
// the node rows, skips rows at which its samples are all zero (their own bitmap) or at which no
// lane's group is all zero, and does acc[s] = fma(z ? len : 0, p_qs, acc[s]) with the 32 values
// of the row read as shared-memory broadcasts (rows are staged 32 at a time with cp.async,
// double buffered, live rows only).
constexpr int SS_WARPS = 4;
constexpr size_t SS_SMEM = (size_t)SS_WARPS * 2 * 32 * 32 * sizeof(double);

__global__ void __launch_bounds__(SS_WARPS * 32)
k_group_skipsums(const double* __restrict__ emb, const double* __restrict__ len_rel,
                 const uint32_t* __restrict__ nz, const int32_t* __restrict__ glist, int n_g,
                 const int32_t* __restrict__ sblist, int64_t npad, int64_t rows, int64_t QR,
                 double* __restrict__ S /*[npad/32][npad]*/) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* buf = reinterpret_cast<double*>(smem_raw) + warp * 2 * 1024;
    const int64_t sb = sblist[blockIdx.x];
    const int64_t words = QR / 32;
    const int gi = blockIdx.y * 32 + lane;
    const int gid = gi < n_g ? glist[gi] : -1;
    const uint32_t* own_row = nz + (4 * sb + warp) * words;
    const uint32_t* z_row = gid >= 0 ? nz + (int64_t)gid * words : nullptr;
    const double* src = emb + (sb * QR) * TILE + warp * 32 + (lane & 15) * 2;
    const int rsel = lane >> 4;  // lanes 0-15 copy even rows, 16-31 odd rows
    const int nwords = (int)((rows + 31) / 32);

    double acc[32];
#pragma unroll
    for (int s = 0; s < 32; ++s) acc[s] = 0.0;

    uint32_t own_next = nwords > 0 ? own_row[0] : 0u;
    {
        double* dst = buf + (lane & 15) * 2;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int r = 2 * i + rsel;
            if ((own_next >> r) & 1u) cp_async16(dst + r * 32, src + (int64_t)r * TILE);
        }
        cp_async_commit();
    }
    for (int w = 0; w < nwords; ++w) {
        const uint32_t own = own_next;
        own_next = w + 1 < nwords ? own_row[w + 1] : 0u;
        {
            double* dst = buf + ((w + 1) & 1) * 1024 + (lane & 15) * 2;
            const double* sp = src + (int64_t)(w + 1) * 32 * TILE;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int r = 2 * i + rsel;
                if ((own_next >> r) & 1u) cp_async16(dst + r * 32, sp + (int64_t)r * TILE);
            }
            cp_async_commit();
        }
        cp_async_wait1();
        __syncwarp();
        if (own != 0u) {
            const uint32_t zw = z_row ? ~z_row[w] : 0u;  // set bit = this lane's group is all zero
            uint32_t todo = own & __reduce_or_sync(0xffffffffu, zw);
            if (todo != 0u) {
                const double lenw = len_rel[(int64_t)w * 32 + lane];
                const double* B = buf + (w & 1) * 1024;
#pragma unroll 1
                while (todo != 0u) {
                    const int b = __ffs(todo) - 1;
                    todo &= todo - 1u;
                    const double lb = __shfl_sync(0xffffffffu, lenw, b);
                    const double zl = ((zw >> b) & 1u) ? lb : 0.0;
                    const double2* row = reinterpret_cast<const double2*>(B + b * 32);
#pragma unroll
                    for (int h = 0; h < 16; ++h) {
                        const double2 v = row[h];
                        acc[2 * h] = fma(zl, v.x, acc[2 * h]);
                        acc[2 * h + 1] = fma(zl, v.y, acc[2 * h + 1]);
                    }
                }
            }
        }
        __syncwarp();
    }
    if (gid >= 0) {
        double* o = S + (int64_t)gid * npad + sb * TILE + warp * 32;
#pragma unroll
        for (int s = 0; s < 32; ++s) o[s] += acc[s];
    }
}

// out += S[g(i)][j] + S[g(j)][i] on the owned rows, before k_finalize applies the epilogue
__global__ void __launch_bounds__(256)
k_add_skipsums(double* __restrict__ out, const int64_t* __restrict__ owned,
               const int64_t* __restrict__ row_base, int64_t n, int64_t npad,
               const double* __restrict__ S) {
    const int64_t b = owned[blockIdx.x / TILE];
    const int64_t i0 = b * TILE, i = i0 + (blockIdx.x % TILE);
    if (i >= n - 1) return;
    const int64_t off = row_base[b] + (cond_row_start(i, n) - cond_row_start(i0, n));
    const int64_t cnt = n - 1 - i;
    const double* si = S + (i / 32) * npad;
    for (int64_t t = threadIdx.x; t < cnt; t += blockDim.x) {
        const int64_t j = i + 1 + t;
        out[off + t] += si[j] + S[(j / 32) * npad + i];
    }
}

// ---------------------------------------------------------------------------------------
// Pair kernel.  Producer, ring and arithmetic loop are k_pair_weighted_v3's: every node row of
// the chunk is staged (18 bulk copies per 16 rows).  Each consumer warp takes the two bitmap
// words of its own row / column group straight from global memory (one word covers two stages,
// the next one is prefetched) and visits only the staged rows that are live for its 32 x 32
// patch; the slot of the next visit is found while the current one computes, so the
// shared-memory loads of an iteration depend on nothing computed in it.
// (Tried: a gathering producer that stages only rows live for the half tile - lane b owns bit
//  b of the bitmap word, slots by prefix popcount, one bulk copy per row and operand.  UBLKCP
//  takes uniform registers, so per-lane copies serialise into an ELECT/R2UR loop of ~9
//  instructions each; 982 ms against 776 ms for this kernel at 10 000 x 100 000 tips.)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(W3_THREADS, 1)
k_pair_weighted_mask(const double* __restrict__ emb, const double* __restrict__ len_q,
                     const uint32_t* __restrict__ nz, unsigned long long* __restrict__ executed,
                     PairArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_base = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)W3_STAGES * W3_STAGE_DOUBLES * sizeof(double));
    uint64_t* empty = full + W3_STAGES;

    const int tid = threadIdx.x;
    const int tile = blockIdx.x >> 1, half = blockIdx.x & 1;
    const int bi = a.tiles[2 * tile], bj = a.tiles[2 * tile + 1];
    const int nstage = (int)((a.qb - a.qa) / W3_KS);

    if (tid == 0) {
        for (int s = 0; s < W3_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], W3_CWARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (tid >= W3_CWARPS * 32) {
        if (tid == W3_CWARPS * 32) {
            const double* Ag = emb + ((int64_t)bi * a.QR + a.qa) * TILE + half * W3_ROWS;
            const double* Bg = emb + ((int64_t)bj * a.QR + a.qa) * TILE;
            const double* Lg = len_q + a.qa;
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < nstage; ++it) {
                if (it >= W3_STAGES) mbar_wait(&empty[s], ph ^ 1u);
                double* A = stage_base + (size_t)s * W3_STAGE_DOUBLES;
                double* B = A + W3_KS * TILE;
                double* L = B + W3_KS * TILE;
                mbar_arrive_expect_tx(&full[s], (uint32_t)((W3_KS * W3_ROWS + W3_KS * TILE + W3_KS) * sizeof(double)));
#pragma unroll
                for (int k = 0; k < W3_KS; ++k)
                    bulk_g2s(A + k * TILE, Ag + ((int64_t)it * W3_KS + k) * TILE, W3_ROWS * sizeof(double), &full[s]);
                bulk_g2s(B, Bg + (int64_t)it * W3_KS * TILE, W3_KS * TILE * sizeof(double), &full[s]);
                bulk_g2s(L, Lg + it * W3_KS, W3_KS * sizeof(double), &full[s]);
                if (++s == W3_STAGES) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
        return;
    }

    const int warp = tid >> 5, lane = tid & 31;
    const int pr = warp >> 2, pc = warp & 3;
    const int ty = lane >> 3, tx = lane & 7;
    const int aoff = pr * 32 + ty * 2;
    const int boff = pc * 32 + tx * 2;
    const int64_t words = a.QR / 32;
    const uint32_t* gA = nz + (int64_t)(4 * bi + 2 * half + pr) * words + a.qa / 32;
    const uint32_t* gB = nz + (int64_t)(4 * bj + pc) * words + a.qa / 32;
    const int nwords = (nstage + 1) / 2;

    double acc[8][4], d[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            acc[r][c] = 0.0;
            d[r][c] = 0.0;
        }
    double lc = 0.0;
    unsigned nexec = 0;

    uint32_t live = 0, live_next = nwords > 0 ? (gA[0] & gB[0]) : 0u;
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nstage; ++it) {
        if ((it & 1) == 0) {
            live = live_next;
            const int wn = (it >> 1) + 1;
            live_next = wn < nwords ? (gA[wn] & gB[wn]) : 0u;
            nexec += __popc(live);
        }
        uint32_t mask = (it & 1) ? (live >> 16) : (live & 0xffffu);
        int k = __ffs(mask) - 1;
        mbar_wait(&full[s], ph);
        const double* S0 = stage_base + (size_t)s * W3_STAGE_DOUBLES;
#pragma unroll 1
        while (k >= 0) {
            const double* P = S0 + k * TILE;
            const double* Lp = S0 + 2 * W3_KS * TILE + k;
            mask &= mask - 1u;
            k = __ffs(mask) - 1;
            double av[8], bv[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const double2 t = *reinterpret_cast<const double2*>(P + aoff + h * 8);
                av[2 * h] = t.x;
                av[2 * h + 1] = t.y;
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const double2 t = *reinterpret_cast<const double2*>(P + W3_KS * TILE + boff + g * 16);
                bv[2 * g] = t.x;
                bv[2 * g + 1] = t.y;
            }
            const double ln = *Lp;
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    acc[r][c] = fma(fabs(d[r][c]), lc, acc[r][c]);
                    d[r][c] = av[r] - bv[c];
                }
            lc = ln;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == W3_STAGES) {
            s = 0;
            ph ^= 1u;
        }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fma(fabs(d[r][c]), lc, acc[r][c]);
    if (executed && lane == 0 && nexec) atomicAdd(executed, (unsigned long long)nexec);

    const int64_t base = a.row_base[bi];
    const int64_t i0 = (int64_t)bi * TILE;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int64_t i = i0 + half * W3_ROWS + aoff + (r >> 1) * 8 + (r & 1);
        if (i >= a.n) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int64_t j = (int64_t)bj * TILE + boff + (c >> 1) * 16 + (c & 1);
            emit(a, base, i0, i, j, acc[r][c]);
        }
    }
}

}  // namespace b2d
