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
//
// S is not summed over the embedding.  The nodes at which a group is all zero form whole
// subtrees (a zero node has only zero descendants), and p_qs is the sum of the sample's entries
// below q, so
//   S[g][s] = sum over the CSR entries (node t, proportion p_ts) of sample s of p_ts * E[g][t],
//   E[g][t] = sum of len over the path from t up to, not including, its first ancestor at which
//             g is non-zero (0 when g is non-zero at t itself)
// E is filled from the roots of the maximal zero subtrees (zero node with a non-zero parent):
// every node x below such a root c gets rd(x) - rd(parent(c)), rd = distance to the root kept
// as a double-double so that the difference is exact to the last bit of E.  Work: one bit test
// per (group, node) + one multiply-add per (group, CSR entry) - no pass over the embedding.
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

// Transposed zero bitmaps: zt[q0 + q][w] bit j = 1 when group gperm[32w + j] is all zero at
// resident row q (0 for padding entries gperm = -1).  One warp per (32 groups, 32 rows): a
// 32 x 32 bit transpose by ballots.
__global__ void __launch_bounds__(256)
k_transpose_bitmaps(const uint32_t* __restrict__ nz, const int32_t* __restrict__ gperm, int64_t ngw,
                    int64_t words, int64_t nwords, int64_t q0, uint32_t* __restrict__ zt /*[mpad][ngw]*/) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= ngw * nwords) return;
    const int64_t gw = warp % ngw, w = warp / ngw;
    const int g = gperm[32 * gw + lane];
    const uint32_t zero_rows = g >= 0 ? ~nz[(int64_t)g * words + w] : 0u;
    uint32_t mine = 0;
#pragma unroll
    for (int b = 0; b < 32; ++b) {
        const uint32_t m = __ballot_sync(0xffffffffu, (zero_rows >> b) & 1u);
        if (lane == b) mine = m;
    }
    zt[(q0 + w * 32 + lane) * ngw + gw] = mine;
}

// The same transposed bitmaps from the bitmaps of the PACKED dense rows (sparse rows handled by the
// intersection kernel, b2d_tips.cuh): zt[pos[r]][w] for the packed rows r, every other row of zt left 0
// ("present").  The rows not packed are whole small subtrees whose branches are left out of the root
// distances the closed form uses (d_rdR: they inherit their parent's value), so a zero subtree rooted
// inside one contributes exactly 0 and its bits are not needed; a zero subtree rooted at a packed row
// covers its position range whatever the bits below say.  One warp per (packed row, 32 groups).
__global__ void __launch_bounds__(256)
k_scatter_zero_bits(const uint32_t* __restrict__ nzd /*[groups][words_d]*/, int64_t words_d,
                    const int32_t* __restrict__ pos, int64_t nd, const int32_t* __restrict__ gperm, int64_t ngw,
                    uint32_t* __restrict__ zt /*[mpad][ngw]*/) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= nd * ngw) return;
    const int64_t r = warp / ngw, gw = warp % ngw;
    const int g = gperm[32 * gw + lane];
    const bool zero = g >= 0 && !((nzd[(int64_t)g * words_d + (r >> 5)] >> (r & 31)) & 1u);
    const uint32_t m = __ballot_sync(0xffffffffu, zero);
    if (lane == 0) zt[(int64_t)pos[r] * ngw + gw] = m;
}

// E[x][gl] for the GW words [kbeg, kbeg + GW) of the permuted group list (gl = 32 * word + lane);
// E is zero-initialised.  One warp per (node q, word): lanes whose group has a maximal zero
// subtree rooted at q fill the subtree's position range [q - size + 1, q].
template <int GW>
__global__ void __launch_bounds__(256)
k_fill_zero_paths(const uint32_t* __restrict__ zt, int64_t ngw, int kbeg, int kcnt,
                  const int32_t* __restrict__ parent_q, const int32_t* __restrict__ size_q,
                  const double* __restrict__ rd_hi, const double* __restrict__ rd_lo, int64_t m,
                  double* __restrict__ E /*[mpad][32 * GW]*/) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= m * kcnt) return;
    const int64_t q = warp / kcnt;
    const int k = (int)(warp % kcnt);
    const uint32_t zq = zt[q * ngw + kbeg + k];
    if (zq == 0u) return;
    const int32_t par = parent_q[q];
    const uint32_t roots = par >= 0 ? (zq & ~zt[(int64_t)par * ngw + kbeg + k]) : zq;
    if (roots == 0u) return;
    const double bh = par >= 0 ? rd_hi[par] : 0.0, bl = par >= 0 ? rd_lo[par] : 0.0;
    const int64_t lo = q - size_q[q] + 1;
    if (q - lo < 32) {  // the usual case: a handful of nodes, one lane per group
        if (!((roots >> lane) & 1u)) return;
        double* e = E + 32 * k + lane;
        for (int64_t x = lo; x <= q; ++x) e[x * (32 * GW)] = (rd_hi[x] - bh) + (rd_lo[x] - bl);
    } else {            // a whole clade without the group: lanes share the position range
        for (int64_t x = lo + lane; x <= q; x += 32) {
            const double v = (rd_hi[x] - bh) + (rd_lo[x] - bl);
            double* e = E + x * (32 * GW) + 32 * k;
            for (uint32_t r = roots; r != 0u; r &= r - 1u) e[__ffs(r) - 1] = v;
        }
    }
}

// S[g][s] = sum over the CSR entries of sample s of p * E[node][g]; one warp per sample, lane =
// group within a word, GW words per thread.  V = int64 counts (divided by the sample total, as
// k_counts_to_proportions does) or double values (features, taken as they are).
template <int GW, typename V>
__global__ void __launch_bounds__(256)
k_sample_pathsums(const int64_t* __restrict__ indptr, const int32_t* __restrict__ qidx,
                  const V* __restrict__ values, const long long* __restrict__ total,
                  const double* __restrict__ E, const int32_t* __restrict__ gperm, int kbeg, int kcnt,
                  const int32_t* __restrict__ sblist, int64_t n, int64_t npad,
                  double* __restrict__ S /*[npad/32][npad]*/) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t s = (int64_t)sblist[blockIdx.x >> 4] * TILE + (blockIdx.x & 15) * 8 + warp;
    if (s >= n) return;
    double acc[GW];
#pragma unroll
    for (int k = 0; k < GW; ++k) acc[k] = 0.0;
    const int64_t e0 = indptr[s], e1 = indptr[s + 1];
    double inv_by = 1.0;
    bool divide = false;
    if (total) {
        const long long t = total[s];
        divide = t > 0;
        inv_by = (double)t;
    }
    for (int64_t eb = e0; eb < e1; eb += 32) {
        const int64_t e = eb + lane;
        int32_t q = 0;
        double pv = 0.0;
        if (e < e1) {
            q = qidx[e];
            pv = values ? (double)values[e] : 1.0;
            if (divide) pv = pv / inv_by;
        }
        const int cnt = (int)min((int64_t)32, e1 - eb);
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const int32_t qj = __shfl_sync(0xffffffffu, q, j);
            const double pj = __shfl_sync(0xffffffffu, pv, j);
            const double* row = E + (int64_t)qj * (32 * GW) + lane;
#pragma unroll
            for (int k = 0; k < GW; ++k)
                if (k < kcnt) acc[k] = fma(pj, row[32 * k], acc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < GW; ++k) {
        if (k >= kcnt) continue;
        const int g = gperm[32 * (kbeg + k) + lane];
        if (g >= 0) S[(int64_t)g * npad + s] = acc[k];
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
// Pair kernel.  Producer and arithmetic loop are k_pair_weighted_v3's: every node row of the
// chunk is staged (34 bulk copies per 32 rows).  Each consumer warp takes the two bitmap
// words of its own row / column group straight from global memory (one word per stage, the
// next one is prefetched) and visits only the staged rows that are live for its 32 x 32
// patch; the slot of the next visit is found while the current one computes, so the
// shared-memory loads of an iteration depend on nothing computed in it.
// (Tried, twice: a gathering producer that stages only rows live for the half tile.  (1) A whole
//  warp, lane b owning bit b of the bitmap word, slots by prefix popcount, one bulk copy per row
//  and operand: UBLKCP takes uniform registers, so per-lane copies serialise into an ELECT/R2UR
//  loop of ~9 instructions each - 982 ms against 776 ms for this kernel at 10 000 x 100 000
//  tips.  (2) One lane walking the set bits, B rows and lengths copied per RUN of live rows,
//  consumers compacting their masks with a lane-parallel bit-extract: 813 ms against 700 ms,
//  and 577 against 349 ms at lambda = 0.005 - the data-dependent FLO/branch chain per run
//  cannot feed eight consumer warps.  Also tried: 64 x 64 quadrants with two CTAs per SM so
//  that one ring's waits are filled by the other's work - 748 ms, the doubled operand traffic
//  and copy count cost more than the waits.)
// ---------------------------------------------------------------------------------------
// Stage = 32 node rows = one bitmap word: A rows compact (64 values), B rows (128 values),
// 32 lengths; 4 stages (193 KB).  (Tried: 8 stages of 16 rows - 3 % slower.)
constexpr int MK_KS = 32;
constexpr int MK_STAGES = 4;
constexpr int MK_STAGE_DOUBLES = MK_KS * W3_ROWS + MK_KS * TILE + MK_KS;
constexpr size_t MK_SMEM = (size_t)MK_STAGES * MK_STAGE_DOUBLES * sizeof(double) + 2 * MK_STAGES * sizeof(uint64_t);

__global__ void __launch_bounds__(W3_THREADS, 1)
k_pair_weighted_mask(const double* __restrict__ emb, const double* __restrict__ len_q,
                     const uint32_t* __restrict__ nz, unsigned long long* __restrict__ executed,
                     PairArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_base = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)MK_STAGES * MK_STAGE_DOUBLES * sizeof(double));
    uint64_t* empty = full + MK_STAGES;

    const int tid = threadIdx.x;
    const int tile = blockIdx.x >> 1, half = blockIdx.x & 1;
    const int bi = a.tiles[2 * tile], bj = a.tiles[2 * tile + 1];
    const int nstage = (int)((a.qb - a.qa) / MK_KS);   // chunks are multiples of 128 rows

    if (tid == 0) {
        for (int s = 0; s < MK_STAGES; ++s) {
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
                if (it >= MK_STAGES) mbar_wait(&empty[s], ph ^ 1u);
                double* A = stage_base + (size_t)s * MK_STAGE_DOUBLES;
                double* B = A + MK_KS * W3_ROWS;
                double* L = B + MK_KS * TILE;
                mbar_arrive_expect_tx(&full[s], (uint32_t)(MK_STAGE_DOUBLES * sizeof(double)));
#pragma unroll 8
                for (int k = 0; k < MK_KS; ++k)
                    bulk_g2s(A + k * W3_ROWS, Ag + ((int64_t)it * MK_KS + k) * TILE, W3_ROWS * sizeof(double), &full[s]);
                bulk_g2s(B, Bg + (int64_t)it * MK_KS * TILE, MK_KS * TILE * sizeof(double), &full[s]);
                bulk_g2s(L, Lg + it * MK_KS, MK_KS * sizeof(double), &full[s]);
                if (++s == MK_STAGES) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
        return;
    }

    const int warp = tid >> 5, lane = tid & 31;
    // the two warps of a sub-partition (warp & 3) take different column groups, so its load is
    // the sum of two less correlated visit counts
    const int pr = warp >> 2, pc = (warp + 2 * pr) & 3;
    const int ty = lane >> 3, tx = lane & 7;
    const int aoff = pr * 32 + ty * 2;
    const int boff = MK_KS * W3_ROWS + pc * 32 + tx * 2;
    const int64_t words = a.QR / 32;
    const uint32_t* gA = nz + (int64_t)(4 * bi + 2 * half + pr) * words + a.qa / 32;
    const uint32_t* gB = nz + (int64_t)(4 * bj + pc) * words + a.qa / 32;

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

    uint32_t live_next = nstage > 0 ? (gA[0] & gB[0]) : 0u;
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nstage; ++it) {
        uint32_t mask = live_next;
        live_next = it + 1 < nstage ? (gA[it + 1] & gB[it + 1]) : 0u;
        nexec += __popc(mask);
        int k = __ffs(mask) - 1;
        mbar_wait(&full[s], ph);
        const double* S0 = stage_base + (size_t)s * MK_STAGE_DOUBLES;
#pragma unroll 1
        while (k >= 0) {
            const double* PA = S0 + k * W3_ROWS + aoff;
            const double* PB = S0 + k * TILE + boff;
            const double* Lp = S0 + MK_KS * (W3_ROWS + TILE) + k;
            mask &= mask - 1u;
            k = __ffs(mask) - 1;
            double av[8], bv[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const double2 t = *reinterpret_cast<const double2*>(PA + h * 8);
                av[2 * h] = t.x;
                av[2 * h + 1] = t.y;
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const double2 t = *reinterpret_cast<const double2*>(PB + g * 16);
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
        if (++s == MK_STAGES) {
            s = 0;
            ph ^= 1u;
        }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fma(fabs(d[r][c]), lc, acc[r][c]);
    if (executed && lane == 0 && nexec) atomicAdd(executed, (unsigned long long)nexec);

    const int boff0 = pc * 32 + tx * 2;
    const int64_t base = a.row_base[bi];
    const int64_t i0 = (int64_t)bi * TILE;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int64_t i = i0 + half * W3_ROWS + aoff + (r >> 1) * 8 + (r & 1);
        if (i >= a.n) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int64_t j = (int64_t)bj * TILE + boff0 + (c >> 1) * 16 + (c & 1);
            emit(a, base, i0, i, j, acc[r][c]);
        }
    }
}

}  // namespace b2d
