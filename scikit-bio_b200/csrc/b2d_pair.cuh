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
    EPI_JACCARD = 4,        // x / ((c_i + c_j + x)/2); denominator 0 -> 0 like scipy
    EPI_BRAYCURTIS_INTERSECT = 5,  // x = sum of minima: (S_i + S_j - 2x) / (S_i + S_j)
    EPI_JACCARD_INTERSECT = 6      // x = |u and v|: (c_i + c_j - 2x) / (c_i + c_j - x); 0/0 -> 0
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
        case EPI_BRAYCURTIS_INTERSECT: {
            const double t = svec[i] + svec[j];
            return (t - 2.0 * x) / t;
        }
        case EPI_JACCARD_INTERSECT: {
            const double t = svec[i] + svec[j];
            const double den = t - x;
            return den == 0.0 ? 0.0 : (t - 2.0 * x) / den;
        }
        default:
            return x;
    }
}

// Add one accumulated value of pair (i, j) (global sample indices) into the raw-sum buffer of
// this launch (store on the buffer's first launch).  The metric's epilogue is applied once at
// the end by k_finalize.
__device__ __forceinline__ void emit(const PairArgs& a, int64_t base, int64_t i0, int64_t i,
                                     int64_t j, double v) {
    if (i >= j || j >= a.n) return;
    int64_t off = base + (cond_index(i, j, a.n) - cond_row_start(i0, a.n));
    a.out[off] = a.first ? v : a.out[off] + v;
}

// out = epilogue(out + out2): one CTA per owned sample row, its condensed entries are
// contiguous.  out2 may be NULL (single launch).
__global__ void __launch_bounds__(256)
k_finalize(double* __restrict__ out, const double* __restrict__ out2,
           const int64_t* __restrict__ owned, const int64_t* __restrict__ row_base, int64_t n,
           int epilogue, const double* __restrict__ svec, const long long* __restrict__ total) {
    const int64_t b = owned[blockIdx.x / TILE];
    const int64_t i0 = b * TILE, i = i0 + (blockIdx.x % TILE);
    if (i >= n - 1) return;
    const int64_t off = row_base[b] + (cond_row_start(i, n) - cond_row_start(i0, n));
    const int64_t cnt = n - 1 - i;
    for (int64_t t = threadIdx.x; t < cnt; t += blockDim.x) {
        double x = out[off + t];
        if (out2) x += out2[off + t];
        out[off + t] = apply_epilogue(x, epilogue, svec, total, i, i + 1 + t);
    }
}

constexpr int WK_KS = 16;  // nodes per staging step of the plain cross-check kernel

// =========================================================================================
// Weighted kernel (default):  acc[i][j] += len_q * |p_q,i - p_q,j|  (also dense Bray-Curtis
// with len == 1), scheduled for instruction-level parallelism instead of relying on many warps.
//   * one CTA owns a 64x128 half tile (grid = 2 x tiles); 8 compute warps, each a 32x32
//     patch of pairs, each thread an 8x4 register tile; thread patches are 2-D inside the warp
//     (4 row groups x 8 column groups) so one node costs 6 shared-memory wavefronts per warp;
//   * the node loop is NOT unrolled and carries the 32 differences d = p_i - p_j of the
//     current node in registers while the DFMAs consume the differences computed one
//     iteration earlier: a DFMA is never issued behind the DADD it depends on (ptxas pairs
//     them back to back whenever both sit in one basic block - ncu round 1: FP64 pipe
//     67-78 % busy, "wait" stalls on the straggling warps);
//   * an FP64 warp instruction holds the sub-partition's dispatch port for 2 cycles, so every
//     other instruction costs FP64 time: the loop body is 64 FP64 + 7 LDS + 7 integer/branch
//     instructions per node and thread (A and B rows share one shared-memory pitch so one
//     pointer addresses both).
// 6-stage ring of {16 x 64 A values (pitch 128), 16 x 128 B values, 16 lengths}, bulk copies.
// (Tried: lengths in the A-row padding, read at the top of the iteration instead of carried in
//  a register - two instructions fewer per node but the exposed load latency cost 7 %.)
// =========================================================================================
constexpr int W3_KS = 16;
constexpr int W3_STAGES = 6;
constexpr int W3_ROWS = 64;
constexpr int W3_CWARPS = 8;
constexpr int W3_THREADS = W3_CWARPS * 32 + 32;
constexpr int W3_STAGE_DOUBLES = 2 * W3_KS * TILE + W3_KS;  // A (pitch 128) + B + len
constexpr size_t W3_SMEM = (size_t)W3_STAGES * W3_STAGE_DOUBLES * sizeof(double) + 2 * W3_STAGES * sizeof(uint64_t);

__global__ void __launch_bounds__(W3_THREADS, 1)
k_pair_weighted_v3(const double* __restrict__ emb, const double* __restrict__ len_q, PairArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_base = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)W3_STAGES * W3_STAGE_DOUBLES * sizeof(double));
    uint64_t* empty = full + W3_STAGES;

    const int tid = threadIdx.x;
    const int tile = blockIdx.x >> 1, half = blockIdx.x & 1;
    const int bi = a.tiles[2 * tile], bj = a.tiles[2 * tile + 1];
    const int64_t nstage = (a.qb - a.qa) / W3_KS;

    if (tid == 0) {
        for (int s = 0; s < W3_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], W3_CWARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (tid >= W3_CWARPS * 32) {
        // ---------------- producer warp ----------------
        if (tid == W3_CWARPS * 32) {
            const double* Ag = emb + ((int64_t)bi * a.QR + a.qa) * TILE + half * W3_ROWS;
            const double* Bg = emb + ((int64_t)bj * a.QR + a.qa) * TILE;
            const double* Lg = len_q + a.qa;
            int s = 0;
            uint32_t ph = 0;
            for (int64_t it = 0; it < nstage; ++it) {
                if (it >= W3_STAGES) mbar_wait(&empty[s], ph ^ 1u);
                double* A = stage_base + (size_t)s * W3_STAGE_DOUBLES;
                double* B = A + W3_KS * TILE;
                double* L = B + W3_KS * TILE;
                mbar_arrive_expect_tx(&full[s], (uint32_t)((W3_KS * W3_ROWS + W3_KS * TILE + W3_KS) * sizeof(double)));
#pragma unroll
                for (int k = 0; k < W3_KS; ++k)
                    bulk_g2s(A + k * TILE, Ag + (it * W3_KS + k) * TILE, W3_ROWS * sizeof(double), &full[s]);
                bulk_g2s(B, Bg + it * W3_KS * TILE, W3_KS * TILE * sizeof(double), &full[s]);
                bulk_g2s(L, Lg + it * W3_KS, W3_KS * sizeof(double), &full[s]);
                if (++s == W3_STAGES) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
        return;
    }

    // ---------------- compute warps ----------------
    const int warp = tid >> 5, lane = tid & 31;
    const int pr = warp >> 2, pc = warp & 3;   // 32x32 patch of the 64x128 half tile
    const int ty = lane >> 3, tx = lane & 7;   // 8x4 pairs per thread
    // rows   pr*32 + h*8 + ty*2 + e   (h = 0..3, e = 0..1)  -> acc row index 2h+e
    // cols   pc*32 + g*16 + tx*2 + e  (g = 0..1, e = 0..1)  -> acc col index 2g+e
    const int aoff = pr * 32 + ty * 2;
    const int boff = pc * 32 + tx * 2;

    double acc[8][4], d[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            acc[r][c] = 0.0;
            d[r][c] = 0.0;
        }
    double lc = 0.0;  // length of the node whose differences are held in d

    int s = 0;
    uint32_t ph = 0;
    for (int64_t it = 0; it < nstage; ++it) {
        mbar_wait(&full[s], ph);
        const double* P = stage_base + (size_t)s * W3_STAGE_DOUBLES;  // A row; B row at +16*TILE
        const double* Lp = P + 2 * W3_KS * TILE;
        const double* Pend = P + W3_KS * TILE;
#pragma unroll 1
        for (; P != Pend; P += TILE, ++Lp) {
            // operands of this node; the DFMAs below still work on the previous node
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

    // ---------------- epilogue ----------------
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
// Unweighted kernel, formulation P (node-packed bits + byte look-up tables):
//   acc[i][j] += LUT_g[ byte_g(u ^ v) ]   for the 16 byte groups g of a 128-node stripe,
// where LUT_g[x] = sum of the branch lengths of the nodes whose bit is set in x (256 entries
// per group, rebuilt per stripe by the CTA itself from the 128 staged lengths).  One 8-byte
// shared-memory gather + one DADD per EIGHT node-pair terms instead of one predicated DADD
// per term.  The 16 tables are interleaved at 8-byte granularity (entry x of table g at
// x*256 + g*8, the other half of each 256-byte row holds the second LUT buffer), and at
// every step the 16 lanes of a half-warp use 16 different tables (lane h takes byte group
// 4*((h/4 + w) & 3) + ((h%4 + b) & 3) at word step w, byte step b), so every 64-bit gather
// is bank-conflict free.  Bound by shared-memory bandwidth: 16 gathers/clk/SM = 128 terms.
// Row/column bits and lengths arrive through a 4-slot ring of bulk copies.
// =========================================================================================
constexpr int UL_STAGES = 4;
constexpr int UL_LUT_BYTES = 256 * 256;  // 256 entries x (2 buffers x 16 tables x 8 B)
constexpr size_t UL_SMEM = UL_LUT_BYTES + UL_STAGES * sizeof(UStage) + UL_STAGES * sizeof(uint64_t);

__device__ __forceinline__ void ul_build_lut(double* lut, int buf, const double* len, int tid) {
    const int g = tid & 15, xh = tid >> 4;
    const double* ln = len + g * 8;
    const double2 p0 = *reinterpret_cast<const double2*>(ln);
    const double2 p1 = *reinterpret_cast<const double2*>(ln + 2);
    const double2 p2 = *reinterpret_cast<const double2*>(ln + 4);
    const double2 p3 = *reinterpret_cast<const double2*>(ln + 6);
    const double l[4] = {p0.x, p0.y, p1.x, p1.y};
    double hs = (xh & 1) ? p2.x : 0.0;
    hs += (xh & 2) ? p2.y : 0.0;
    hs += (xh & 4) ? p3.x : 0.0;
    hs += (xh & 8) ? p3.y : 0.0;
    double lo[16];
    lo[0] = 0.0;
#pragma unroll
    for (int x = 1; x < 16; ++x) {
        const int low = x & (x - 1);          // x without its lowest set bit
        const int bit = (x & 1) ? 0 : (x & 2) ? 1 : (x & 4) ? 2 : 3;
        lo[x] = lo[low] + l[bit];
    }
    double* dst = lut + (size_t)(xh * 16) * 32 + buf * 16 + g;
#pragma unroll
    for (int xl = 0; xl < 16; ++xl) dst[xl * 32] = hs + lo[xl];
}

__global__ void __launch_bounds__(256, 1)
k_pair_unweighted_lut(const uint4* __restrict__ bits, const double* __restrict__ len_q, PairArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* lut = reinterpret_cast<double*>(smem_raw);
    UStage* st = reinterpret_cast<UStage*>(smem_raw + UL_LUT_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + UL_LUT_BYTES + UL_STAGES * sizeof(UStage));

    const int tid = threadIdx.x;
    const int bi = a.tiles[2 * blockIdx.x], bj = a.tiles[2 * blockIdx.x + 1];
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t s0 = a.qa / NODE_STRIPE;
    const int nst = (int)(a.qb / NODE_STRIPE - s0);
    const uint4* Rg = bits + ((int64_t)bi * a.NS + s0) * TILE;
    const uint4* Cg = bits + ((int64_t)bj * a.NS + s0) * TILE;
    const double* Lg = len_q + s0 * NODE_STRIPE;

    auto issue = [&](int i) {  // stripe i of this launch -> ring slot i % UL_STAGES
        UStage* d = st + (i % UL_STAGES);
        uint64_t* bar = full + (i % UL_STAGES);
        mbar_arrive_expect_tx(bar, (uint32_t)sizeof(UStage));
        bulk_g2s(d->rows, Rg + (int64_t)i * TILE, TILE * sizeof(uint4), bar);
        bulk_g2s(d->cols, Cg + (int64_t)i * TILE, TILE * sizeof(uint4), bar);
        bulk_g2s(d->len, Lg + (int64_t)i * NODE_STRIPE, NODE_STRIPE * sizeof(double), bar);
    };

    if (tid == 0) {
        for (int s = 0; s < UL_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0)
        for (int i = 0; i < nst && i < UL_STAGES - 1; ++i) issue(i);

    double acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.0;

    if (nst > 0) {
        mbar_wait(&full[0], 0);
        ul_build_lut(lut, 0, st[0].len, tid);
    }
    const int hq = tx >> 2, hr = tx & 3;
    const char* lut_bytes = reinterpret_cast<const char*>(lut);

    for (int i = 0; i < nst; ++i) {
        __syncthreads();  // LUT[i&1] complete; LUT[(i+1)&1] and ring slot (i+3)%4 are free
        if (tid == 0 && i + UL_STAGES - 1 < nst) issue(i + UL_STAGES - 1);
        if (i + 1 < nst) {
            mbar_wait(&full[(i + 1) % UL_STAGES], (uint32_t)(((i + 1) / UL_STAGES) & 1));
            ul_build_lut(lut, (i + 1) & 1, st[(i + 1) % UL_STAGES].len, tid);
        }
        mbar_wait(&full[i % UL_STAGES], (uint32_t)((i / UL_STAGES) & 1));
        const uint32_t* R = reinterpret_cast<const uint32_t*>(st[i % UL_STAGES].rows);
        const uint32_t* C = reinterpret_cast<const uint32_t*>(st[i % UL_STAGES].cols);
        const uint32_t bufoff = (uint32_t)(i & 1) * 128u;
#pragma unroll 1
        for (int wt = 0; wt < 4; ++wt) {
            const int wsel = (hq + wt) & 3;
            const uint32_t rot = (uint32_t)hr * 8u;
            uint32_t rw[8], cw[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                uint32_t x = R[(ty * 8 + e) * 4 + wsel];
                uint32_t y = C[(e * 16 + tx) * 4 + wsel];
                rw[e] = __funnelshift_r(x, x, rot);
                cw[e] = __funnelshift_r(y, y, rot);
            }
            // table offsets (bytes within a 256-byte LUT row) for byte steps 0..3
            uint32_t o[4];
#pragma unroll
            for (int bt = 0; bt < 4; ++bt) o[bt] = 8u * (4u * wsel + ((hr + bt) & 3)) + bufoff;
            const uint32_t offA = o[0] | (o[1] << 8), offB = o[2] | (o[3] << 8);
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t X = rw[r] ^ cw[c];
                    const uint32_t a0 = __byte_perm(X, offA, 0x7604);
                    const uint32_t a1 = __byte_perm(X, offA, 0x7615);
                    const uint32_t a2 = __byte_perm(X, offB, 0x7624);
                    const uint32_t a3 = __byte_perm(X, offB, 0x7635);
                    double v = acc[r][c];
                    v += *reinterpret_cast<const double*>(lut_bytes + a0);
                    v += *reinterpret_cast<const double*>(lut_bytes + a1);
                    v += *reinterpret_cast<const double*>(lut_bytes + a2);
                    v += *reinterpret_cast<const double*>(lut_bytes + a3);
                    acc[r][c] = v;
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
__global__ void __launch_bounds__(256, 1)
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
